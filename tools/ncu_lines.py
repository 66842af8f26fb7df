"""Joins an ncu report's per-SASS-address counters with nvdisasm line info.
   python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel-substring> [topN]
Prints instructions executed and stall samples per CUDA source line, and the opcode mix."""
import collections, csv, io, os, re, subprocess, sys, tempfile

def main():
    rep, so, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # split per kernel
    blocks = []; cur = None
    for r in rows:
        if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; blocks.append(cur)
        elif cur is not None: cur["rows"].append(r)
    blk = [b for b in blocks if kname in b["name"]][0]
    hdr = blk["rows"][0]
    ia, isrc, ie, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    addr = {}
    base = None
    for r in blk["rows"][1:]:
        if len(r) < len(hdr): continue
        a = int(r[ia], 16)
        if base is None: base = a
        addr[a - base] = (int(r[ie]), int(r[isamp]), r[isrc])
    td = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cub)], capture_output=True, text=True).stdout
    # locate the kernel's text section by mangled-name fragments of kname
    line_of = {}; cur_line = None; inside = False; files = {}
    mang = None
    for ln in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
        if m:
            inside = all(tok in m.group(1) for tok in re.findall(r"[A-Za-z_]+|\d+", kname.replace("(int)", "").replace("(bool)", "")) if tok not in ("void", "int"))
            continue
        if not inside: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2))); files[os.path.basename(m.group(1))] = m.group(1); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m: line_of[int(m.group(1), 16)] = cur_line
    per_line = collections.Counter(); samp_line = collections.Counter(); ops = collections.Counter(); tot = 0
    per_line_ops = collections.defaultdict(collections.Counter)
    for a, (n, s, src) in addr.items():
        l = line_of.get(a, ("?", -1))
        per_line[l] += n; samp_line[l] += s; tot += n
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src); op = m.group(2) if m else "?"
        ops[op] += n; per_line_ops[l][op] += n
    texts = {b: (open(f, errors="replace").read().splitlines() if os.path.exists(f) else []) for b, f in files.items()}
    print(f"kernel {blk['name']}: {tot} warp instructions, {sum(samp_line.values())} samples")
    for l, n in per_line.most_common(top):
        text = texts.get(l[0], [])
        code = text[l[1] - 1].strip()[:90] if 0 < l[1] <= len(text) else ""
        mix = " ".join(f"{o}:{c*100//max(n,1)}%" for o, c in per_line_ops[l].most_common(4))
        print(f"{n:11d} {100*n/tot:5.1f}% smp {100*samp_line[l]/max(1,sum(samp_line.values())):5.1f}%  {l[0]}:{l[1]:<4d} {code}\n{'':30s}[{mix}]")
    print("opcode mix:", ", ".join(f"{o} {100*c/tot:.1f}%" for o, c in ops.most_common(24)))

main()
