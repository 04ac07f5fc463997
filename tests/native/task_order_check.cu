// Host-only checker of the factorisation kernel's claim lists (hetero.cuh / flags.cuh): the deadlock-freedom argument of
// chol_hetero_tma_kernel rests on ONE property — workers claim tasks in list order and a task only waits for tasks that
// appear EARLIER in the joint order — so this program rebuilds the lists for a range of shapes and verifies it.
// Compiled with nvcc (the lists live in CUDA headers) and run on the CPU by tests/test_host_logic.py; exit code 0 = all shapes hold.
#include <cstdio>
#include <map>
#include <tuple>
#include <vector>
#include "../../gsum_b200/csrc/hetero.cuh"

typedef std::tuple<int, int, int> Key;          // (i, k, b)

static int check_many(int T, int Trows, int batch, bool thin_last, int delay) {
    std::vector<int4> all;
    df_build_tasks(all, T, Trows, batch, false, thin_last, delay);
    std::map<Key, int> pos;
    for (int p = 0; p < (int)all.size(); p++) {
        Key k(all[p].x, all[p].y, all[p].z);
        if (pos.count(k)) { printf("duplicate task (%d,%d,%d)\n", all[p].x, all[p].y, all[p].z); return 1; }
        pos[k] = p;
    }
    // every tile of the bordered lower triangle exactly once
    if ((int64_t)all.size() != (int64_t)batch * ((int64_t)T * (T + 1) / 2 + (int64_t)(Trows - T) * T)) { printf("task count\n"); return 1; }
    auto before = [&](int i, int k, int b, int p) { auto it = pos.find(Key(i, k, b)); return it != pos.end() && it->second < p; };
    for (int p = 0; p < (int)all.size(); p++) {
        const int i = all[p].x, k = all[p].y, b = all[p].z;
        if (i < k || i >= Trows || k >= T || b < 0 || b >= batch) { printf("bad task\n"); return 1; }
        if (i == k) { if (k > 0 && !before(k, k - 1, b, p)) { printf("diag (%d,%d,%d) before its last operand\n", i, k, b); return 1; } }
        else {
            if (!before(k, k, b, p)) { printf("panel (%d,%d,%d) before M_kk\n", i, k, b); return 1; }
            if (k > 0 && (!before(i, k - 1, b, p) || !before(k, k - 1, b, p))) { printf("panel (%d,%d,%d) before its operands\n", i, k, b); return 1; }
        }
        const bool thin = (all[p].w & 1) != 0;
        if (thin != (thin_last && i == Trows - 1 && i >= T)) { printf("thin flag\n"); return 1; }
    }
    // the split keeps the order inside each list and puts the column-0 diagonal tiles at the head of the factor list
    std::vector<int4> gemm, fact;
    ht_build_tasks(gemm, fact, T, Trows, batch, false, thin_last, delay);
    if ((int)fact.size() != batch * T) { printf("factor list size\n"); return 1; }
    for (int p = 0; p < batch; p++) if (fact[p].x != 0) { printf("column-0 tiles not first\n"); return 1; }
    int last = -1;
    for (const int4 &t : gemm) { int q = pos[Key(t.x, t.y, t.z)]; if (q <= last) { printf("gemm list reordered\n"); return 1; } last = q; }
    last = -1;
    for (const int4 &t : fact) { int q = pos[Key(t.x, t.x, t.y)]; if (q <= last) { printf("factor list reordered\n"); return 1; } last = q; }
    return 0;
}

// chain mode: band tiles (k, k) and (k+1, k) belong to the chain CTAs (always resident, one per matrix); the list holds the panel
// tasks of rows >= k+2 and of the border, and the `pre` tasks (flag bit 1) that prepare (k+1, k) and (k+1, k+1) without their last two terms
static int check_chain(int T, int Trows, int batch, bool thin_last) {
    std::vector<int4> g;
    ht_build_chain_tasks(g, T, Trows, batch, thin_last);
    std::map<Key, int> pos, prepos;
    for (int p = 0; p < (int)g.size(); p++) {
        Key k(g[p].x, g[p].y, g[p].z);
        std::map<Key, int> &m = (g[p].w & 2) ? prepos : pos;
        if (m.count(k)) { printf("chain: duplicate\n"); return 1; }
        m[k] = p;
    }
    auto final_before = [&](int i, int j, int b, int p) {      // tile (i, j) final before list position p?  band tiles: the chain CTA's job
        if (i == j || i == j + 1) return true;                  // produced by the chain CTA of matrix b, which is never parked
        auto it = pos.find(Key(i, j, b));
        return it != pos.end() && it->second < p;
    };
    for (int p = 0; p < (int)g.size(); p++) {
        const int i = g[p].x, k = g[p].y, b = g[p].z;
        const bool pre = (g[p].w & 2) != 0;
        if (!pre) {
            if (!(i >= k + 2 || i >= T)) { printf("chain: band tile (%d,%d) in the panel list\n", i, k); return 1; }
            if (k > 0 && (!final_before(i, k - 1, b, p) || !final_before(k, k - 1, b, p))) { printf("chain: panel (%d,%d,%d) before its operands\n", i, k, b); return 1; }
        } else {
            const int nj = (i == k) ? k - 2 : k - 1;            // terms a pre task applies: j < nj
            if (nj > 0 && (!final_before(i, nj - 1, b, p) || !final_before(k, nj - 1, b, p))) { printf("chain: pre (%d,%d,%d) before its operands\n", i, k, b); return 1; }
        }
    }
    // every non-band tile once; pre tasks exactly for the band steps k -> k+1 with k >= 2
    int64_t want = 0;
    for (int k = 0; k < T; k++) want += (int64_t)batch * ((Trows - (k + 1 < T ? k + 2 : k + 1)));
    if ((int64_t)pos.size() != want) { printf("chain: panel count %zu vs %lld\n", pos.size(), (long long)want); return 1; }
    if ((int64_t)prepos.size() != (int64_t)2 * batch * (T > 3 ? T - 3 : 0)) { printf("chain: pre count\n"); return 1; }
    return 0;
}

int main() {
    int bad = 0, shapes = 0;
    for (int T : {1, 2, 3, 4, 7, 16})
        for (int nb : {0, 1, 3})
            for (int batch : {1, 2, 5})
                for (int thin = 0; thin < 2; thin++)
                    for (int delay : {0, 3, 1000}) {
                        if (thin && nb == 0) continue;
                        bad += check_many(T, T + nb, batch, thin != 0, delay);
                        if (delay == 0) bad += check_chain(T, T + nb, batch, thin != 0);
                        shapes++;
                    }
    printf("%d shapes checked, %d violations\n", shapes, bad);
    return bad ? 1 : 0;
}
