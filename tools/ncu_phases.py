#!/usr/bin/env python3
"""Where a BVH trace kernel spends its warp instructions, by PHASE of a path segment (ncu capture + nvdisasm line info).

  cuobjdump -xelf all rt_b200/lib/librtcu.so && nvdisasm --print-line-info-inline rtcu.sm_100a.cubin > dis.txt
  ncu -i rep.ncu-rep --page source --csv > src.csv
  python tools/ncu_phases.py src.csv dis.txt <mangled-kernel-substring> [samples traced by the launch] [--list "<phase>"]

tools/ncu_lines.py charges an instruction to the INNERMOST source line of its inline chain, which lumps every packed FP32
intrinsic, shuffle and load into the header it comes from.  Here the whole chain (`File ..., line N inlined at ...`) is used: an
instruction belongs to the outermost phase function that appears anywhere in its chain -- trav_step / slab_pair /
bvh_leaf_pair_test, beam_closest_sphere, generate, shade_segment (split into Philox, normal, material, scatter, sky), the rest is
the kernel's own loop.  Function line ranges are read from the sources named in the line info, so the script follows the code --
which must then be the code of the captured build (`--sources DIR`: read kernels.cuh / spec.cuh from DIR instead, e.g. a
`git worktree` of the commit the capture names).
Output: warp-instruction share, stall-sample share, active lanes and thread instructions per traced sample of each phase.
"""
import collections, csv, re, sys

def function_ranges(path):
    """{name: (first line, last line)} of the top-level functions of a source file"""
    out, text = {}, open(path).read().split("\n")
    heads = [i for i, l in enumerate(text) if re.match(r"(template|__device__|__global__|__host__|struct|constexpr|#|//|namespace|static)", l)]
    for i, l in enumerate(text):
        m = re.match(r"(?:__device__|__global__|__host__)[^(]*?\b(\w+)\s*\(", re.sub(r"__launch_bounds__\([^)]*\)", "", l))
        rest = "\n".join(text[i:i + 12])
        if m and "{" in rest and (";" not in rest or rest.index("{") < rest.index(";")):  # a definition, not a forward declaration
            nxt = next((h for h in heads if h > i and text[h].startswith(("template", "__device__", "__global__", "__host__", "struct", "constexpr", "// ----"))), len(text))
            out.setdefault(m.group(1), (i + 1, nxt))
    return out



def phase(chain, K, S):
    def has(tab, fname, *names):
        return any(f == fname and any(tab[n][0] <= ln <= tab[n][1] for n in names if n in tab) for f, ln in chain)
    k = lambda *n: has(K, "kernels.cuh", *n)
    s = lambda *n: has(S, "spec.cuh", *n)
    philox = s("philox4x32", "rng_block", "philox_keys")
    if k("beam_closest_sphere"):
        return "primary list scan: leaf pair tests" if k("bvh_leaf_pair_test", "bvh_leaf_candidate") else "primary list scan: loop, loads"
    if k("trav_step", "closest_sphere_bvh"):
        if k("bvh_leaf_pair_test", "bvh_leaf_candidate"):
            return "traversal: leaf pair tests"
        if k("slab_pair"):
            return "traversal: slab tests"
        if k("trav_init"):
            return "traversal: init"
        return "traversal: node loads, child bookkeeping, stack"
    if k("trav_init"):
        return "traversal: init"
    if k("closest_hit_linear"):  # scan scenes
        if s("sphere_candidate"):
            return "scan: sphere candidates (sqrt, t, acceptance)"
        if s("sphere_pair_test"):
            return "scan: packed pair tests"
        if s("plane_test"):
            return "scan: planes"
        return "scan: loop, loads, select"
    if k("generate"):
        return "generate: Philox" if philox else "generate: primary ray"
    if k("shade_segment"):
        if philox:
            return "shade: Philox"
        if k("hit_normal_global", "hit_normal"):
            return "shade: normal"
        if k("load_material", "hit_material"):
            return "shade: material"
        if s("scatter", "schlick", "random_unit_vector", "reflect3", "scatter_kind"):
            return "shade: scatter"
        if s("sky"):
            return "shade: sky"
        return "shade: other"
    if k("combine_with_planes", "closest_hit_bvh", "closest_plane_cold", "closest_hit_scan_cold"):
        return "combine with planes / fallback"
    return "kernel loop: claim, ballots, epilogue"



def main():
    skip = {i + 1 for i, a in enumerate(sys.argv) if a in ("--list", "--sources")}
    args = [a for i, a in enumerate(sys.argv) if i >= 1 and i not in skip and not a.startswith("--")]
    src_csv, dis_txt, kernel_sub = args[:3]
    n_samples = float(args[3]) if len(args) > 3 else None
    want = sys.argv[sys.argv.index("--list") + 1] if "--list" in sys.argv else None

    rows = list(csv.reader(open(src_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hdr_i + 1:] if len(r) >= len(hdr) and r[0].startswith("0x")]
    base = int(data[0][0], 16)
    ex = {int(r[0], 16) - base: (int(r[ci["Instructions Executed"]]), int(r[ci["Thread Instructions Executed"]]), int(r[ci["# Samples"]]), r[ci["Source"]].strip())
          for r in data}

    # ---- inline chains per instruction offset
    lines = open(dis_txt).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kernel_sub in l and l.rstrip().endswith(":"))
    chains, cur, done, last, files = {}, [], [], [], {}
    for l in lines[start + 1:]:
        if l.startswith("//---") or l.startswith("\t.section"):
            break
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            files[m.group(1).split("/")[-1]] = m.group(1)
            cur.append((m.group(1).split("/")[-1], int(m.group(2))))
            if "inlined at" not in m.group(3):
                done.append(cur)
                cur = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", l)
        if m:
            if done:
                last, done = done[-1], []
            chains[int(m.group(1), 16)] = last


    src_dir = sys.argv[sys.argv.index("--sources") + 1] if "--sources" in sys.argv else None
    K = function_ranges(f"{src_dir}/kernels.cuh" if src_dir else files["kernels.cuh"])
    S = function_ranges(f"{src_dir}/spec.cuh" if src_dir else files["spec.cuh"])


    agg, tot = collections.defaultdict(lambda: [0, 0, 0]), [0, 0, 0]
    for off, (wi, ti, sm, sass) in ex.items():
        a = agg[phase(chains.get(off, []), K, S)]
        for j, v in enumerate((wi, ti, sm)):
            a[j] += v
            tot[j] += v
    print(f"{kernel_sub}: warp instructions {tot[0]:.4g}, thread instructions {tot[1]:.4g}, active lanes {tot[1] / tot[0]:.2f}" +
          (f", thread instructions per sample {tot[1] / n_samples:.0f}" if n_samples else ""))
    for c, (wi, ti, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"  {c:50s} warp instr {100 * wi / tot[0]:5.1f} %   stall samples {100 * sm / max(tot[2], 1):5.1f} %   lanes {ti / max(wi, 1):5.1f}" +
              (f"   thread instr / sample {ti / n_samples:7.1f}" if n_samples else ""))
    if want:
        print("---- instructions of phase:", want)
        for off in sorted(ex):
            wi, ti, sm, sass = ex[off]
            ch = chains.get(off, [])
            if phase(ch, K, S) == want and wi > tot[0] * 2e-4:
                print(f"{off:6x} {100 * wi / tot[0]:5.2f}% lanes {ti / max(wi, 1):5.1f} stalls {100 * sm / max(tot[2], 1):4.2f}%  {sass[:72]:72s} {ch[0][0]}:{ch[0][1]}" if ch else sass)


if __name__ == "__main__":
    main()
