"""Opcode histogram of the hot kernels' SASS (cuobjdump -sass of the built library): the evidence that the tensor-core linears are
tcgen05 (UTC*MMA, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier) and the vector
stages warp-level HMMA.   python tools/sass_histogram.py > profiles/<tag>_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "flowmol_b200", "libflowmol_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
want = re.compile(r"UTC\w*MMA|LDTM|STTM|UBLKCP[\w.]*|UTCBAR[\w.]*|SYNCS[\w.]*|HMMA[\w.]*|UCGABAR\w*|UTCATOMSWS|STG\.E\.ENL2\.256|LDGSTS[\w.]*|MUFU\.\w+")
hist, name = {}, None
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"fm::Dims<[^>]*>", "D", name).split("(")[0].replace("void fm::", "")
        hist[name] = collections.Counter()
        continue
    if name:
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][\w.]*)", line)
        if m:
            hist[name]["total"] += 1
            w = want.match(m.group(1))
            if w:
                hist[name][w.group(0)] += 1
keep = [k for k in hist if re.search(r"k_egemm_(p|e|g|g2|c)<|k_vecr|k_vec_[abc]|k_ctmc|k_decode|k_node_(pre|mid)|k_conv_edge|k_edge_(init|head)_r", k)]
for k in sorted(keep):
    c = hist[k]
    print(f"{k}: {c['total']} instructions; " + ", ".join(f"{op} x{n}" for op, n in sorted(c.items()) if op != "total"))
