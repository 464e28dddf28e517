#!/usr/bin/env python
"""SASS instruction summary per kernel of libdcora_b200.so (cuobjdump -sass / -res-usage):
registers, shared memory, instruction count and the mnemonics that matter for the roofline
argument (DFMA/DMUL/DADD, global/shared loads and stores, TMA bulk copies, mbarrier waits,
barriers, shuffles).  python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dcora_b200", "libdcora_b200.so")
FILT = subprocess.run(["c++filt"], input="", capture_output=True, text=True).returncode == 0


def demangle(n):
    if not FILT:
        return n
    return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()


def short(n):
    n = demangle(n).replace("(anonymous namespace)::", "").replace("dcora::", "")
    n = re.sub(r"^void\s+", "", n)
    return re.sub(r"\(.*$", "", n)


GROUPS = [("DFMA", r"^DFMA"), ("DMUL/DADD", r"^(DMUL|DADD)"), ("MUFU", r"^MUFU"),
          ("LDG", r"^LDG"), ("STG", r"^STG"), ("LDS", r"^LDS"), ("STS", r"^STS"),
          ("LDC/ULDC", r"^(LDC|ULDC)"), ("TMA(UBLKCP)", r"^UBLKCP"), ("L2PF(UBLKPF)", r"^UBLKPF"),
          ("SYNCS(mbar)", r"^SYNCS"), ("BAR", r"^BAR"), ("UCGABAR(cluster)", r"^UCGABAR"),
          ("SHFL", r"^SHFL"), ("ATOM/RED", r"^(ATOM|RED|ATOMG|ATOMS)"), ("BRA", r"^BRA"),
          ("IMAD/IADD", r"^(IMAD|IADD3|IADD|LEA)")]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = (int(m.group(1)), int(m.group(2)))
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            kernels[cur][m.group(1).split(".")[0]] += 1
    want = sys.argv[1:] or ["trsv_tape_kernel<5, 4, 4, 16>", "trsv_tape_kernel<5, 4, 1, 8>",
                            "bsr_spmm_kernel<5, 4>", "rem_spmm_kernel<5>", "manifold_kernel<5, 4,",
                            "dots_kernel", "axpby_kernel", "tcg_update_kernel", "coupling", "rtr_cluster_kernel<5, 4, 1, 16>",
                            "round_kernel", "gram_kernel", "lanczos", "certify"]
    print(f"{'kernel':58s} {'regs':>4s} {'smem':>6s} {'instr':>6s}  " +
          " ".join(f"{g[0]:>8s}" for g in GROUPS[:14]))
    for name, cnt in kernels.items():
        s = short(name)
        if not any(w in s for w in want):
            continue
        total = sum(cnt.values())
        reg, sh = usage.get(name, (0, 0))
        cols = []
        for g, pat in GROUPS[:14]:
            cols.append(sum(v for k, v in cnt.items() if re.match(pat, k)))
        print(f"{s[:58]:58s} {reg:4d} {sh:6d} {total:6d}  " + " ".join(f"{c:8d}" for c in cols))
    print("\ncolumns: static SASS instruction counts (not executed counts); smem = static shared "
          "memory in bytes (the K5 ring is dynamic shared memory: 8 slots x record bytes per warp).")


if __name__ == "__main__":
    main()
