"""profiles/sass_k2_r02.txt: instruction mix of the grid kernels from `cuobjdump -sass` of the built object (no GPU needed)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "ennemi_b200", "_build", "eb2_ksg2.o")
WANT = ("knn_kernel2<4>", "leftover_kernel2<4>", "count_psi_kernel", "layout_kernel", "bucket_scatter_kernel", "bucket_hist_kernel",
        "fine_cells_kernel", "sample_rank_kernel", "knn3_kernel<4, 8>", "leftover3_kernel<4, 8>", "layout3_kernel", "knn1d_kernel")
out = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True).stdout
names = subprocess.run(["bash", "-c", f"cuobjdump -sass {OBJ} | grep 'Function :' | sed 's/.*Function : //' | c++filt"], capture_output=True, text=True).stdout.split("\n")
blocks = out.split("Function : ")[1:]
with open(os.path.join(ROOT, "profiles", "sass_k2_r02.txt"), "w") as f:
    f.write("# SASS of the grid kernels (cuobjdump -sass ennemi_b200/_build/eb2_ksg2.o, sm_100a), round 2, final build (tools/sass_summary.py)\n"
            "# No tensor-core / TMA instructions by design: the searches read a few dozen to a few hundred candidates per row straight from\n"
            "# L1/L2 (LDG.E.64), compare with DADD + DSETP (|q - c| < thr per coordinate) and keep a sorted top-(k+1) list in registers\n"
            "# (DSETP + FSEL/SEL).  The TMA bulk copies (UBLKCP) live in the general path's kernels (profiles/sass_knn_hotloop_r01.txt).\n\n")
    for name, blk in zip(names, blocks):
        if not any(w in name for w in WANT):
            continue
        ops = collections.Counter()
        for line in blk.split("\n"):
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
            if m:
                ops[m.group(1)] += 1
        tot = sum(ops.values())
        f.write(f"## {name}\ninstructions: {tot}\nmix: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(18)) + "\n\n")
print("wrote profiles/sass_k2_r02.txt")
