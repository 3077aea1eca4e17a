"""Static SASS instruction mix of the hot kernels of libmm2b200.so (cuobjdump -sass; no GPU needed).
usage: python tools/sass_mix.py > profiles/rNN_sass_mix.md"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "minimap2_rs_b200", "libmm2b200.so")
WANT = ["sketch_tile_kernel_v4<unsigned int, 10, false>", "sketch_tile_kernel_v4<unsigned long, 10, false>", "sketch_tile_kernel_v4<unsigned long, 10, true>",
        "chain_ring_kernel<false, true>", "chain_ring_kernel<false, false>", "chain_dense_kernel<16>", "seed_hits_kernel<0>",
        "anchor_msort_kernel<128, true, false>", "os_pass_kernel", "unpack_reads_kernel", "lookup_build_kernel", "tab_fill_kernel"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
fn, mix = None, collections.defaultdict(collections.Counter)
for line in out.split("\n"):
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::|void |\(.*$", "", fn)
        fn = re.sub(r"\((bool|int)\)", "", fn)
        fn = fn.replace("<0", "<false").replace(", 0>", ", false>").replace(", 1>", ", true>").replace("<1", "<true") if "chain_ring" in fn else fn
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        mix[fn][m.group(1)] += 1
print("# SASS of libmm2b200.so (cuobjdump -sass): instruction mix of the hot kernels (`python tools/sass_mix.py`)\n")
print("Cubins in the library: %s.  Static counts, rare paths included.  Nothing on this path is a dense contraction, so there are no" % ", ".join(archs))
print("tensor-core (UTCMMA / HMMA) or TMA (UBLKCP / UTMALDG) instructions; the movers are 128-bit LDG / STG (256-bit `LDG.E.ENL2.256` for")
print("the seed-lookup table line), window extrema use VIMNMX3, warp reductions REDUX, MATCH.ANY ranks the radix digits.\n")
print("| kernel | instructions | top mnemonics |\n|---|---|---|")
for w in WANT:
    key = [k for k in mix if k.replace("(anonymous namespace)::", "") == w or k == w]
    if not key:
        key = [k for k in mix if w.split("<")[0] in k and w.split("<")[1][:-1].replace(" ", "") in k.replace(" ", "")]
    if not key:
        print("| `%s` | not found | |" % w)
        continue
    c = mix[key[0]]
    print("| `%s` | %d | %s |" % (key[0], sum(c.values()), ", ".join("%s %d" % kv for kv in c.most_common(16))))
