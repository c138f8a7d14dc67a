#!/usr/bin/env python3
"""Attribute the warp-stall samples of one kernel in an ncu report to source lines.
The report's SASS page gives samples per instruction; `nvdisasm -g` of the cubin (built with -lineinfo) gives the
source line of every instruction (the innermost inlined frame); joined by instruction offset.
usage: python profiles/sass_hotspots.py <report.ncu-rep> <kernel regex> <cubin> <mangled-name regex> [top N]
"""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, kre, cubin, fre = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][idx["Address"]], 16)
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    line_of, cur, infn = {}, None, False
    for l in dis.splitlines():
        if l.startswith(".text."):
            infn = re.search(fre, l) is not None
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: collections.Counter())
    tot = inst = 0
    for r in data:
        off = int(r[idx["Address"]], 16) - base
        key = line_of.get(off, ("?", 0))
        s, n = int(r[idx["# Samples"]]), int(r[idx["Instructions Executed"]])
        a = agg[key]
        a["samples"] += s
        a["inst"] += n
        for h in stalls:
            a[h] += int(r[idx[h]])
        tot += s
        inst += n
    print(f"kernel {rows[0][1]}: {tot} samples, {inst} warp instructions, {len(data)} SASS instructions")
    srcs = {}
    for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        if f not in srcs:
            try:
                p = subprocess.run(["find", "/root/repo/mars_ode_physics_b200/csrc", "-name", f], capture_output=True, text=True).stdout.split()[0]
                srcs[f] = open(p).read().splitlines()
            except Exception:
                srcs[f] = []
        text = srcs[f][ln - 1].strip()[:70] if 0 < ln <= len(srcs[f]) else ""
        st = " ".join(f"{h[6:]}={a[h]}" for h in sorted(stalls, key=lambda h: -a[h])[:3] if a[h])
        print(f"{100 * a['samples'] / tot:5.1f}% smp {100 * a['inst'] / max(inst, 1):5.1f}% inst  {f}:{ln:<4} {text:<70} {st}")


if __name__ == "__main__":
    main()
