#!/usr/bin/env python
"""Warp-stall samples of one kernel of an ncu report, grouped into code regions (runs of SASS instructions with the same executed
count = the roles of a warp-specialised kernel: producer / MMA issuer / split warps / epilogue / waits).

    python tools/stall_regions.py report.ncu-rep kernel_regex > profiles/rNN_<kernel>_stall_regions.txt
"""
import collections
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hi]
    ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[hi + 1:]:
        if len(r) <= iex or not r[ia].startswith("0x"):
            continue
        data.append((int(r[ia], 16), r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0), {h: int(r[hdr.index(h)] or 0) for h in st}))
    base = data[0][0]
    total = sum(d[2] for d in data)
    print("# %s: %d SASS instructions, %d warp-stall samples" % (rows[0][1] if len(rows[0]) > 1 else kern, len(data), total))
    print("# region = consecutive instructions with (about) the same executed count; columns: address range, executed count per")
    print("# instruction, samples (share), instructions, top stall reasons, first instruction")
    seg, cur = [], None
    for a, s, n, e, sd in data:
        if cur is None or (e != cur["e"] and abs(e - cur["e"]) > 0.2 * max(e, cur["e"], 1)):
            cur = {"a0": a - base, "a1": a - base, "e": e, "n": 0, "cnt": 0, "st": collections.Counter(), "first": s}
            seg.append(cur)
        cur["a1"] = a - base
        cur["n"] += n
        cur["cnt"] += 1
        cur["st"].update(sd)
    for s in seg:
        if s["n"] * 200 >= total:
            top = ", ".join("%s %d" % (k.replace("stall_", ""), v) for k, v in s["st"].most_common(4) if v)
            print("%#7x-%#7x  exec %8d  samples %5d (%4.1f%%)  %4d instr | %s | %s" % (s["a0"], s["a1"], s["e"], s["n"], 100.0 * s["n"] / total, s["cnt"], top, s["first"][:48]))


if __name__ == "__main__":
    main()
