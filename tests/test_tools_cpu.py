"""tools/ncu_phases.py follows the sources: the phase functions it attributes instructions to must be found, as definitions."""
import importlib.util
import pathlib

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_phase_functions_are_found_in_the_sources():
    spec = importlib.util.spec_from_file_location("ncu_phases", ROOT / "tools" / "ncu_phases.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    K = mod.function_ranges(str(ROOT / "rt_b200" / "csrc" / "kernels.cuh"))
    S = mod.function_ranges(str(ROOT / "rt_b200" / "csrc" / "spec.cuh"))
    text = (ROOT / "rt_b200" / "csrc" / "kernels.cuh").read_text().split("\n")
    for name in ("closest_hit_linear", "bvh_leaf_candidate", "bvh_leaf_pair_test", "trav_init", "slab_pair", "trav_step", "closest_sphere_bvh",
                 "beam_closest_sphere", "combine_with_planes", "closest_hit_bvh", "hit_normal_global", "hit_material", "load_material", "generate",
                 "shade_segment", "k_render_stragglers"):
        lo, hi = K[name]
        assert lo < hi and name in text[lo - 1], name
    # shade_segment is declared before segment_step and defined after it: the range must be the definition's
    assert K["shade_segment"][0] > K["segment_step"][0]
    for name in ("philox4x32", "rng_block", "sphere_candidate", "sphere_pair_test", "plane_test", "sky", "scatter", "schlick", "random_unit_vector"):
        lo, hi = S[name]
        assert lo < hi, name
    # a chain through trav_step and slab_pair is a slab test; through shade_segment and philox a shade Philox
    mid = lambda r: (r[0] + r[1]) // 2
    assert mod.phase([("sm_100_rt.hpp", 108), ("kernels.cuh", mid(K["slab_pair"])), ("kernels.cuh", mid(K["trav_step"]))], K, S) == "traversal: slab tests"
    assert mod.phase([("spec.cuh", mid(S["philox4x32"])), ("kernels.cuh", mid(K["shade_segment"]))], K, S) == "shade: Philox"
    assert mod.phase([("kernels.cuh", mid(K["k_render_stragglers"]))], K, S) == "kernel loop: claim, ballots, epilogue"
