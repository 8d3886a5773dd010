"""Parity at the BASELINE configurations the benchmark runs (C2 128^3 f32, C3 256^3 u16, C4 512^3 f32): the workloads are built
exactly as bench.py builds them and every output of the per-iteration hot path is compared with the flat oracle (all host
threads; bit-identical to the unmodified reference on the fixtures, tests/test_oracle_golden.py).

Integer decisions (sample -> voxel classification, sample counts, z-ray voxel counts, Gauss sets through the forces) are compared
bit for bit; floating-point sums to 1e-10 (a single sample in a wrong voxel moves a per-label total by > 1e-8 relative)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
THREADS = max(1, os.cpu_count() or 1)


def _workload(pkg, name):
    import bench

    def factory(vol, nb_phase):
        img = pkg.Image3D(vol)
        return img, pkg.SegmentFunction(img, nb_phase, 0.01)

    vol, cfg, edge, snap, img, seg = bench.build_workload(pkg, name, factory, 0.2)
    seg.set_snapshot(snap)
    return {"vol": vol, "cfg": cfg, "snap": snap, "img": img, "seg": seg, "default_variant": seg.get_tet_variant()}


@pytest.fixture(scope="module", params=["C2", "C3", "C4"])
def work(request, pkg, flat):
    name = request.param
    w = _workload(pkg, name)
    vol, snap = w["vol"], w["snap"]
    ref = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], c=[0.0], threads=THREADS)
    w["ref"] = ref
    w["name"] = name
    yield w
    w["img"].close()


def test_labels_are_the_oracles(work, pkg, flat):
    """bench.py labels the tets from the GPU's own per-tet means: they must be the labels the oracle's means give."""
    syn = pkg.synthetic
    cfg = syn.CONFIGS[work["name"]]
    snap, ref = work["snap"], work["ref"]
    base = np.where(snap["tet_label"] != 0, 1, 0)
    assert np.array_equal(syn.labels_from_means(ref["mean"], base, cfg["thresholds"]), snap["tet_label"])


def test_sample_classification(work):
    seg, ref = work["seg"], work["ref"]
    h, ns = seg.tet_sample_hash()
    assert np.array_equal(ns, ref["nsamp"])
    assert np.array_equal(h, ref["hash"])


@pytest.mark.parametrize("variant", [None, 131072 | 128 | 16384])
def test_hot_path_sums(work, pkg, flat, variant):
    """The instantiation the benchmark times: per-label sums only (no per-tet outputs)."""
    import torch
    seg, ref, snap = work["seg"], work["ref"], work["snap"]
    P = work["cfg"]["nb_phase"]
    seg.set_tet_variant(work["default_variant"] if variant is None else variant)
    try:
        lab = snap["tet_label"]
        tot, volu, mean = flat.phase_reduce(lab, ref["total"], ref["vol"], P)
        sq, _, _ = flat.phase_reduce(lab, ref["sq"][:, 0], ref["vol"], P)
        for want_sq in (False, True):
            seg.set_hot_want_sq(want_sq)
            d_sums = torch.zeros(3 * P, dtype=torch.float64, device="cuda")
            seg.tet_phase_sums_dev(d_sums.data_ptr())
            work["img"].synchronize()
            s = d_sums.cpu().numpy()
            assert seg.last_sample_count() == int(ref["nsamp"][lab != 0].sum())
            np.testing.assert_allclose(s[:P], tot, rtol=1e-10)
            np.testing.assert_allclose(s[2 * P:], volu, rtol=1e-10)
            if want_sq:
                np.testing.assert_allclose(s[P:2 * P], sq, rtol=1e-9)
        # per-tet outputs of the same kernel family: volumes bit for bit, totals bit for bit when the voxels are floats
        o = seg.tet_integrate(include_bound=True)
        assert np.array_equal(o["vol"], ref["vol"])
        if work["vol"].dtype == np.float32:
            assert np.array_equal(o["total"], ref["total"])
        else:
            np.testing.assert_allclose(o["total"], ref["total"], rtol=1e-13, atol=1e-14)
    finally:
        seg.set_hot_want_sq(False)
        seg.set_tet_variant(work["default_variant"])


def test_image_force_eval(work, pkg, flat):
    """dscgpu_image_force_eval, the call the adapter and bench.py's e2e leg make: tet-sampled means, then the external forces."""
    seg, ref, snap, vol = work["seg"], work["ref"], work["snap"], work["vol"]
    P = work["cfg"]["nb_phase"]
    mean, F = seg.image_force_eval(snap, mean_mode=pkg.MEAN_TET, force_mode=pkg.FORCE_GATHER)
    _, _, mean_ref = flat.phase_reduce(snap["tet_label"], ref["total"], ref["vol"], P)
    np.testing.assert_allclose(mean, mean_ref, rtol=1e-10)
    F_ref = flat.face_forces(vol, snap, P, mean)  # one thread: the reference's order of additions
    assert np.array_equal(F, F_ref), float(np.abs(F - F_ref).max())
    Fa = seg.compute_external_force(mode=pkg.FORCE_ATOMIC, mean=mean)
    np.testing.assert_allclose(Fa, F_ref, rtol=1e-6, atol=1e-9)


def test_zray_means_and_energy(work, pkg, flat):
    """update_average_intensity / compute_energy at full size: voxel counts bit for bit."""
    seg, snap, vol = work["seg"], work["snap"], work["vol"]
    P = work["cfg"]["nb_phase"]
    work["img"].build_sum_table()
    z = flat.zray(vol, snap, P, mode=0)
    m = seg.update_average_intensity()
    assert np.array_equal(seg._phase_volume, z["count"])
    np.testing.assert_allclose(seg._total_intensities, z["total"], rtol=1e-9)
    np.testing.assert_allclose(m, z["mean"], rtol=1e-9)
    e = flat.zray(vol, snap, P, mode=1, mean=m, alpha=0.01)
    d, a = seg.compute_energy(mean=m)
    np.testing.assert_allclose([d, a], [e["energy_data"], e["energy_area"]], rtol=1e-6)
