#!/usr/bin/env python
"""SASS evidence for the hot kernels: for the all-sky-with-aerosols Float32 kernels of
rrtmgp.jl_b200/csrc/{fast_lw_ng1,fast_sw_ng1,ws_lw_ng1}.o writes profiles/<tag>_sass_<kernel>.txt with
  * registers / spills (from the ptxas log), total instructions, counts of the Blackwell-specific mnemonics
    (LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, SYNCS = mbarrier, UTCBAR / USETMAXREG ...),
  * the main level loop (the basic block between the first gather of the loop and its backward branch) verbatim,
    with its instruction mix.
usage: python profiles/dump_sass.py r2"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "rrtmgp.jl_b200", "csrc")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
KERNELS = [("fast_lw_ng1", "ILi0ELi256ELi1ELb1ELb1ELb0ELi12E", "lw_2stream_fused"),
           ("fast_sw_ng1", "ILi2ELi224ELi1ELb1ELb1ELb0ELi12E", "sw_2stream_fused"),
           ("ws_lw_ng1", "solve_kernel_wsILi0ELi256ELi1ELb1ELb1ELb0E", "lw_2stream_warp_specialised"),
           ("solver_tm", "solve_kernel_tmILi0E", "lw_2stream_float64_tensor_memory")]
SPECIAL = ("LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "USETMAXREG", "UTCBAR", "UTCATOMSWS", "LDG", "LDS", "STS", "MUFU", "FFMA", "FMUL", "FADD")

for obj, key, name in KERNELS:
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(CSRC, obj + ".o")], capture_output=True, text=True).stdout
    fn = [f for f in sass.split("Function :") if key in f.split("\n")[0]]
    if not fn:
        continue
    f = fn[0]
    ins = []
    for l in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?)\s*/\*", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    def opcode(i):
        t = i.split()
        return (t[1] if t[0].startswith("@") else t[0]).split(".")[0].rstrip(";")
    counts = collections.Counter(opcode(i) for _, i in ins)
    # main loop: the largest backward branch region that contains >= 12 LDG
    best = None
    for a, i in ins:
        if "BRA" in i:
            m = re.search(r"0x([0-9a-f]+)", i)
            if m and int(m.group(1), 16) < a:
                t = int(m.group(1), 16)
                body = [(x, y) for x, y in ins if t <= x <= a]
                n_ldg = sum(1 for _, y in body if y.startswith("LDG") or " LDG" in y)
                if n_ldg >= (12 if "solver_tm" not in obj else 20) and (best is None or len(body) < len(best)):
                    best = body
    log = open(os.path.join(CSRC, obj + ".ptxas.log")).read()
    regs = re.findall(r"Used (\d+) registers", log)
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_{name}.txt"), "w") as o:
        o.write(f"# {name}: {f.splitlines()[0].strip()}\n# from rrtmgp.jl_b200/csrc/{obj}.o (cuobjdump -sass), sm_100a\n")
        o.write(f"# instructions in the kernel: {len(ins)}; registers of the object's kernels (ptxas): {sorted(set(map(int, regs)))}\n")
        o.write("# mnemonic counts in the whole kernel: " + ", ".join(f"{k} {counts[k]}" for k in SPECIAL if counts[k]) + "\n")
        if best:
            mix = collections.Counter(opcode(i) for _, i in best)
            o.write(f"# main level loop: {len(best)} instructions per (layer, 32 g-points) step"
                    + (" [4 layers per pass in the warp-specialised gas loop]" if "ws_" in obj else "") + "\n")
            o.write("# mix: " + ", ".join(f"{k} {v}" for k, v in mix.most_common()) + "\n\n")
            for a, i in best:
                o.write(f"/*{a:05x}*/ {i}\n")
    print(name, len(ins), len(best) if best else None)
