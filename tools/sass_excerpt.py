#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel the resource usage, counts of the mnemonics that matter (256-bit global loads /
stores with an L2 policy, bulk copies, local-memory traffic, barriers, atomics) and the listing of its streaming loops.
    python tools/sass_excerpt.py > profiles/r02_sass_excerpt.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "maxstyle_b200", "libmaxstyle_b200.so")
KERNELS = [("fwd_pair_kernel<float, 8, 4, 4>", "fwd_pair_kernelIfLi8ELi4ELi4"), ("bwd_nchw_kernel<float, 8, 256, 2, true>", "bwd_nchw_kernelIfLi8ELi256ELi2ELb1"),
           ("fwd_resident_kernel<__nv_bfloat16, 256, 4>", "fwd_resident_kernelI13__nv_bfloat16Li256ELi4")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines()
    blocks = re.split(r"\n\s*Function : ", sass)
    print("cuobjdump -sass / -res-usage of maxstyle_b200/libmaxstyle_b200.so (sm_100a), hot kernels of the benchmarked step\n")
    for title, key in KERNELS:
        blk = next((b for b in blocks if b.split("\n", 1)[0].find(key) >= 0), None)
        if blk is None:
            print(f"== {title}: not found\n")
            continue
        ins = [l for l in blk.splitlines() if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", l)]
        text = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in ins]
        usage = next((res[i + 1].strip() for i, l in enumerate(res) if key in l and i + 1 < len(res)), "")
        print(f"== {title}\n   {usage}")
        # the main kernel ends at the first EXIT-terminated region; out-of-line helpers (the gap code of the paired forward) follow
        last_exit = max((i for i, l in enumerate(text) if " EXIT" in l and i < len(text)), default=len(text))
        first_ret = next((i for i, l in enumerate(text) if re.search(r"\bRET\b", l)), len(text))
        count = lambda pat, lo=0, hi=None: sum(1 for l in text[lo:hi] if re.search(pat, l))
        main_hi = min(first_ret, len(text))
        for name, pat in (("LDG.E...256 with L2 policy", r"LDG\.E\.\S*ENL2\.256"), ("STG.E...256 with L2 policy", r"STG\.E\.\S*ENL2\.256"),
                          ("UBLKCP (cp.async.bulk)", r"UBLKCP"), ("SYNCS (mbarrier)", r"SYNCS"), ("BAR.SYNC", r"BAR\.SYNC"),
                          ("ATOMG / RED", r"\b(ATOMG|RED)\b"), ("LDL (local loads, whole function incl. out-of-line gap code)", r"\bLDL"),
                          ("STL (local stores, whole function)", r"\bSTL")):
            print(f"   {name:62s} {count(pat):4d}")
        # streaming loops: backward branches whose body holds 256-bit global accesses
        addr = {}
        for i, l in enumerate(text):
            m = re.search(r"/\*([0-9a-f]{4,5})\*/", l)
            if m:
                addr[int(m.group(1), 16)] = i
        shown = 0
        for i, l in enumerate(text):
            m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", l)
            if not m:
                continue
            tgt = addr.get(int(m.group(1), 16))
            if tgt is None or tgt >= i:
                continue
            body = text[tgt:i + 1]
            n256 = sum(1 for b in body if re.search(r"\.256", b))
            if n256 < 2 or len(body) > 4000:
                continue
            loc = sum(1 for b in body if re.search(r"\b(LDL|STL)\b", b))
            print(f"\n   -- streaming loop at {text[tgt].split('*/')[0].split('/*')[-1]} .. {l.split('*/')[0].split('/*')[-1]}: {len(body)} instructions, "
                  f"{n256} 256-bit global accesses, {loc} local-memory accesses")
            keep = [b for b in body if re.search(r"LDG|STG|BRA|BAR|LDL|STL|FFMA|FADD|FMUL", b)]
            for b in keep[:10]:
                print("      " + b.strip()[:150])
            if len(keep) > 10:
                ff = sum(1 for b in keep if re.search(r"FFMA|FADD|FMUL", b))
                print(f"      ... ({len(keep) - 10} more of these; {ff} FP32 arithmetic instructions in the body)")
            shown += 1
            if shown >= 6:
                break
        print()


if __name__ == "__main__":
    main()
