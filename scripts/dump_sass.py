#!/usr/bin/env python
"""SASS evidence for the stepping kernels (north_star: "each choice is evidenced by an ncu capture and SASS listing").

For every kernel named in KERNELS: disassemble it from the shipped library with `cuobjdump -sass`, find its hot loop (the
innermost backward branch that encloses the most FP64 instructions), write the loop to profiles/sass/<name>.sass with a header
that counts its instructions by pipe and divides by the cells one trip of the loop processes (cells per lane x unroll factor).

    python scripts/dump_sass.py            # after the library has been built (no GPU needed)
"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cryogrid.jl_b200", "lib", "libcgb200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")

# file stem -> (demangled prefix, cells per lane, steps per loop trip (source-level unroll factor), note)
KERNELS = {
    "k_heat_euler_fast_7_3": ("void cgb::k_heat_euler_fast<7, 3, true, false>(", 7, 3, "FreeWater fast path, BASELINE cfg2 (N = 218)"),
    "k_heat_sfcc_fast_7": ("void cgb::k_heat_sfcc_fast<7, 384, false, false>(", 7, 2, "SFCC single-class fast path (presolver LUT + curve table in shared memory), cfg4'"),
    "k_heat_sfcc_fast_7_n2": ("void cgb::k_heat_sfcc_fast<7, 384, false, true>(", 7, 2, "SFCC single-class fast path, van Genuchten n == 2 (rsqrt closed form), cfg4''"),
    "k_heat_step_7_1": ("void cgb::k_heat_step<7, 1>(", 7, 1, "general SFCC presolver-LUT kernel (several soil classes); the body holds both the MaxDelta and the CFL variant of stage A"),
    "k_heat_step_7_4": ("void cgb::k_heat_step<7, 4>(", 7, 1, "general SFCC presolver-LUT kernel, every van Genuchten n == 2; the body holds both the MaxDelta and the CFL variant of stage A"),
    "k_heat_step_7_2": ("void cgb::k_heat_step<7, 2>(", 7, 1, "SFCC on-device Newton kernel, BASELINE cfg4 (per-column parameters)"),
    "k_heat_newton_tab_7": ("void cgb::k_heat_newton_tab<7, 8, false>(", 7, 1, "SFCC Newton kernel on per-column curve tables, BASELINE cfg4 (per-column parameters); the loop holds pass 1 (branch-free) and the seven pass-2 bodies, of which ~1.6 run per step"),
    "k_lite_warp_9_3_2": ("void cgb::k_lite_warp<9, 3, 2, false>(", 9, 1, "LiteImplicitEuler warp-per-column kernel, BASELINE cfg3 (N = 278); loop = one Picard iteration"),
    "k_salt_euler_6": ("void cgb::k_salt_euler<6, false>(", 6, 1, "coupled heat + salt kernel, BASELINE cfg5 (N = 178)"),
}

FP64 = re.compile(r"^(DFMA|DMUL|DADD|DSETP|DMNMX|F2F\.F64|I2F\.F64|F2I.*F64|MUFU\.(RCP64H|RSQ64H))")
INSTR = re.compile(r"^\s*/\*([0-9a-f]{4,})\*/\s+(.*?);")


def classify(op):
    o = op.split()[0]
    if o.startswith("@"):
        o = op.split()[1]
    if FP64.match(o):
        return "fp64 pipe"
    if o.startswith("MUFU"):
        return "mufu (other)"
    if o.startswith(("LDS", "STS")):
        return "shared ld/st"
    if o.startswith(("LDG", "STG", "LD.", "ST.", "LDC", "ATOM", "RED", "LDL", "STL")):
        return "global/local/const ld/st"
    if o.startswith(("SHFL", "REDUX", "VOTE", "MATCH")):
        return "shuffle/redux/vote"
    if o.startswith(("BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "BAR", "NOP", "YIELD")):
        return "control"
    return "int/alu/select/move"


def functions():
    txt = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
    chunks = re.split(r"\n\s*Function : \S+\n", "\n" + txt)[1:]
    return dict(zip(names, chunks))


def hot_loop(lines):
    """lines: [(addr, text)].  Innermost backward-branch range with the most FP64 instructions per instruction of nesting."""
    addr_index = {a: i for i, (a, _) in enumerate(lines)}
    loops = []
    for i, (a, t) in enumerate(lines):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:[\w!]+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*(0x[0-9a-f]+)", t) or re.search(r"\bBRA.*?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_index:
                loops.append((addr_index[tgt], i))
    # per-column set-up loops (the table fits of k_heat_newton_tab / k_salt_euler) hold a lot of FP64 work but are not step loops:
    # a step loop exchanges neighbours or reduces the limiter across the warp
    def has_warp_op(lo, hi):
        return any(re.match(r"(@!?U?P\d+\s+)?(SHFL|REDUX|CREDUX)", t) for _, t in lines[lo:hi + 1])
    warp_loops = [l for l in loops if has_warp_op(*l)]
    if warp_loops:
        # several back edges to one header (an unrolled loop) are ONE loop: keep the widest range per header
        by_head = {}
        for lo, hi in warp_loops:
            by_head[lo] = max(by_head.get(lo, hi), hi)
        merged = sorted(by_head.items())
        fp = {l: sum(1 for _, t in lines[l[0]:l[1] + 1] if classify(t) == "fp64 pipe") for l in merged}
        # innermost: contains no other warp-op loop with a fair amount of FP64 work; of those, the one with the most FP64
        inner = [l for l in merged if not any(o != l and l[0] <= o[0] and o[1] <= l[1] and fp[o] >= 20 for o in merged)]
        inner = [l for l in inner if fp[l] >= 20] or inner
        return max(inner, key=lambda l: fp[l])
    best, score = None, -1
    for lo, hi in loops:
        nfp = sum(1 for _, t in lines[lo:hi + 1] if classify(t) == "fp64 pipe")
        inner = [1 for l2, h2 in loops if (l2, h2) != (lo, hi) and lo <= l2 and h2 <= hi and
                 sum(1 for _, t in lines[l2:h2 + 1] if classify(t) == "fp64 pipe") > 0.5 * nfp]
        if inner:
            continue                      # a nested loop holds most of the FP64 work: that one is the hot loop
        if nfp > score:
            best, score = (lo, hi), nfp
    return best


def main():
    os.makedirs(OUT, exist_ok=True)
    funcs = functions()
    summary = []
    for stem, (prefix, cpl, unroll, note) in KERNELS.items():
        match = [k for k in funcs if k.startswith(prefix)]
        if not match:
            print(f"{stem}: not found in {LIB}", file=sys.stderr)
            continue
        lines = [(int(m.group(1), 16), m.group(2).strip()) for m in (INSTR.match(l) for l in funcs[match[0]].split("\n")) if m]
        loop = hot_loop(lines)
        if loop is None:
            print(f"{stem}: no loop found", file=sys.stderr)
            continue
        body = lines[loop[0]:loop[1] + 1]
        cnt = Counter(classify(t) for _, t in body)
        total = len(body)
        cells = cpl * unroll
        mufu = Counter(t.split()[0] if not t.startswith("@") else t.split()[1] for _, t in body if "MUFU" in t)
        hdr = [f"// {match[0]}", f"// {note}", f"// hot loop: {len(body)} SASS instructions at 0x{body[0][0]:x}..0x{body[-1][0]:x} of {len(lines)} in the kernel; "
               f"one trip = {unroll} step(s) x {cpl} cells per lane = {cells} cell-steps per lane",
               "// STATIC instruction counts of the loop body per trip / per cell-step, by pipe (rarely taken paths inside the loop -- saves,",
               "// forcing-cursor advance, bisection of crowded LUT buckets, the CFL variant -- are included; the dynamic counts are in the ncu",
               "// summaries under profiles/):"]
        for k, v in sorted(cnt.items(), key=lambda kv: -kv[1]):
            hdr.append(f"//   {k:28s} {v:5d}   {v / cells:7.2f}")
        hdr.append(f"//   {'TOTAL':28s} {total:5d}   {total / cells:7.2f}")
        hdr.append("// MUFU by kind: " + ", ".join(f"{k} x{v}" for k, v in mufu.items()) if mufu else "// MUFU: none")
        hdr.append(f"// generated by scripts/dump_sass.py from {os.path.relpath(LIB, ROOT)} (cuobjdump -sass)")
        with open(os.path.join(OUT, stem + ".sass"), "w") as f:
            f.write("\n".join(hdr) + "\n")
            for a, t in body:
                f.write(f"        /*{a:04x}*/  {t} ;\n")
        summary.append((stem, total, cells, cnt.get("fp64 pipe", 0)))
        print(f"{stem}: loop {total} instr, {cells} cell-steps per trip -> {total / cells:.1f} instr, {cnt.get('fp64 pipe', 0) / cells:.1f} FP64 per cell-step")
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("# SASS hot loops of the stepping kernels\n\nGenerated by `python scripts/dump_sass.py` (cuobjdump -sass of the shipped "
                "`libcgb200.so`); one file per kernel with per-pipe instruction counts per cell-step in its header.\n\n"
                "| kernel | loop instructions | cell-steps per trip (per lane) | instr / cell-step | FP64-pipe instr / cell-step |\n|---|---|---|---|---|\n")
        for stem, total, cells, fp in summary:
            f.write(f"| `{stem}` | {total} | {cells} | {total / cells:.1f} | {fp / cells:.1f} |\n")


if __name__ == "__main__":
    main()
