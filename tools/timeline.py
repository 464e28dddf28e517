#!/usr/bin/env python
"""GPU timeline of RBCD iterations on the bench workload (CUPTI through torch.profiler):
kernel busy time vs idle gaps, gap histogram, and which kernels precede the long gaps.
    python tools/timeline.py [--steps 2] > profiles/...txt"""
import argparse, collections, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def short(name):
    """kernel name without namespaces / argument list (template arguments kept)."""
    n = name.replace("(anonymous namespace)::", "").replace("dcora::", "")
    n = re.sub(r"^void\s+", "", n)
    m = re.match(r"([\w:]+(<[^(]*>)?)", n)
    return (m.group(1) if m else n)[:44]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=100)
    ap.add_argument("--agents", type=int, default=256)
    ap.add_argument("--rank", type=int, default=5)
    ap.add_argument("--range-fraction", type=float, default=0.1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--restart", action="store_true",
                    help="profile the first iterations from the initial guess (what bench.py times)")
    a = ap.parse_args()
    import torch
    from torch.profiler import profile, ProfilerActivity
    import dcora_b200 as dc
    from dcora_b200 import rbcd
    log = lambda m: print("[timeline] " + m, file=sys.stderr, flush=True)
    meas, gt, idx, gp = bench.build_problem(a, log)
    S = rbcd.RBCD(gp, a.rank, params=dc.ROptParameters(time_bound_s=0.0))
    X0 = bench.initial_X(a, meas, idx, gp)
    S.set_X(X0)
    for _ in range(3):
        S.iterate(); S.monitor()
    if a.restart:
        S.set_X(X0)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(a.steps):
            S.iterate(); S.monitor()
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
            ev.append((e.time_range.start, e.time_range.end, e.name))
    ev.sort()
    if not ev:
        print("no CUDA events captured")
        return
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    span = (t1 - t0) / 1e3
    busy = 0.0
    cur_end = ev[0][0]
    gaps = []
    for s, e, n in ev:
        if s > cur_end:
            gaps.append((s - cur_end, prev_name, n))
        busy += max(0.0, e - max(s, cur_end))
        if e > cur_end:
            cur_end, prev_name = e, n
    print(f"steps {a.steps}: span {span:.2f} ms ({span / a.steps:.2f} ms/step), GPU busy {busy / 1e3:.2f} ms "
          f"({100 * busy / 1e3 / span:.1f} %), {len(ev)} device activities, {len(gaps)} gaps")
    bins = [(0, 2), (2, 5), (5, 10), (10, 20), (20, 50), (50, 100), (100, 1e9)]
    for lo, hi in bins:
        g = [x[0] for x in gaps if lo <= x[0] < hi]
        print(f"  gaps {lo:>4}-{hi if hi < 1e8 else 'inf':>4} us: n={len(g):5d} total={sum(g) / 1e3:7.2f} ms")
    by = collections.defaultdict(lambda: [0, 0.0])
    for d, p, n in gaps:
        k = (short(p), short(n))
        by[k][0] += 1; by[k][1] += d
    print("  largest gap contributors (kernel before -> kernel after):")
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"    {v[1] / 1e3:7.2f} ms  n={v[0]:4d}  avg={v[1] / v[0]:7.1f} us   {k[0]} -> {k[1]}")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for s, e, n in ev:
        agg[short(n)][0] += 1; agg[short(n)][1] += e - s
    print("  device time by activity:")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"    {v[1] / 1e3:7.2f} ms  n={v[0]:5d}  avg={v[1] / v[0]:7.1f} us  {k}")


if __name__ == "__main__":
    main()
