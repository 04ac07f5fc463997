"""Dev tool: overlap analysis of the phase trace of one GEMM CTA (GSUM_B200_DF_STATS=1 GSUM_B200_TRACE_CTA=<cta>).
Reads the '[trace] q n t_claim t_acc t_loopend t_mkk t_done t_pub ik b' lines of the LAST launch in the log and prints, per
group, the task table and how much of the time 0, 1, 2 or 3 groups were inside their DMMA main loop."""
import sys
import numpy as np
lines = [l.split() for l in open(sys.argv[1]) if l.startswith("[trace]")]
# keep the last launch: traces are printed per launch in (q, n) order
runs, cur, last = [], [], (-1, -1)
for l in lines:
    key = (int(l[1]), int(l[2]))
    if key < last and key == (0, 0) or (cur and key <= last and key[0] == 0 and key[1] == 0):
        runs.append(cur); cur = []
    cur.append([int(x) for x in l[1:]]); last = key
runs.append(cur)
tr = np.array(runs[-1], dtype=np.int64)
t0 = tr[:, 2].min()
ev = []
for q in range(3):
    rows = tr[tr[:, 0] == q]
    print(f"group {q}: {len(rows)} tasks")
    for r in rows[: int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
        n, c, a, le, mk, dn, pb, ik, b = r[1:]
        i, k = ik // 1000, ik % 1000
        print(f"  n={n:2d} ({i:2d},{k:2d},b={b:3d}) claim {c - t0:8d} | wait C {a - c:6d} | loop {le - a:7d} ({(le - a) / max(k, 1):6.0f}/step) | "
              f"wait Mkk {(mk - le) if mk else 0:6d} | trsm+store {(dn - (mk if mk else le)):6d} | publish {pb - dn:5d}")
    for r in rows:
        ev.append((r[3], +1)); ev.append((r[4], -1))       # in-loop interval [t_acc, t_loopend]
ev.sort()
tend = tr[:, 7].max()
hist = np.zeros(4); cur = 0; prev = t0
for t, d in ev:
    hist[cur] += t - prev; prev = t; cur += d
hist[cur] += tend - prev
print("time with n groups in the main loop:", " ".join(f"{n}: {100 * h / hist.sum():.1f}%" for n, h in enumerate(hist)), f"| span {tend - t0} cycles")
