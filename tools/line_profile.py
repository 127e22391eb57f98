"""Executed warp instructions per CUDA source line of one kernel: the SASS-level counts of an `ncu --set full
--import-source on` report joined with `nvdisasm -g` line info of the library (built with -lineinfo).

    python tools/line_profile.py gpurun_out/prof.ncu-rep <kernel name substring> [units] [top]

`units` divides the counts (e.g. the env count of the launch) to print instructions per unit."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "isaacgym_b200", "_lib", "libppk.so")


def main():
    rep, pat = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp(prefix="ppk_cubin_")
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
    line_of, name = {}, None
    for cubin in sorted(os.listdir(tmp)):
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        cur, inside = None, False
        for l in dis.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
            if m:
                inside = pat in m.group(1) and not line_of
                name = m.group(1) if inside else name
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", l)
            if m:
                line_of[int(m.group(1), 16)] = cur
        if line_of:
            break
    assert line_of, f"no kernel matching {pat!r} in {LIB}"
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # the report may hold several kernels / launches: take the first block whose kernel name matches
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and pat.split("ILi")[0].split("ILb")[0][-24:] in "".join(r))
    hdr = rows[start + 1]
    ia, iad = hdr.index("Instructions Executed"), hdr.index("Address")
    data = []
    for r in rows[start + 2:]:
        if len(r) <= ia or not r[iad].startswith("0x"):
            break
        data.append(r)
    base = int(data[0][iad], 16)
    agg = collections.Counter()
    for r in data:
        agg[line_of.get(int(r[iad], 16) - base, ("?", 0))] += int(r[ia])
    tot = sum(agg.values())
    print(f"{name}: {tot} warp instructions, {tot / units:.1f} per unit")
    cache = {}
    for (f, ln), c in agg.most_common(top):
        p = os.path.join(ROOT, "isaacgym_b200", "csrc", f)
        if f not in cache:
            cache[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = cache[f][ln - 1].strip()[:90] if 0 < ln <= len(cache[f]) else ""
        print(f"{100 * c / tot:5.1f}% {c / units:8.2f}  {f}:{ln}  {text}")


if __name__ == "__main__":
    main()
