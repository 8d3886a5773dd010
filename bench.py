#!/usr/bin/env python3
"""Benchmark of the DSC image-force evaluation (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4] [--impl reference]

A step = one full image-force evaluation of the workload: per-tet voxel integration + per-phase
means + interface-face Gauss-quadrature forces.  Metric = (tets + interface faces)
processed per second, whole job.

  value      device-timed (CUDA events on the launching stream, max over ranks), mesh snapshot and volume
             already resident in HBM; L2 is flushed between timed steps.
  e2e        same evaluation through the C ABI with HOST buffers: pinned snapshot -> H2D -> kernels ->
             D2H of the per-phase statistics and the vertex forces, every step.
  roofline   dominant kernel (per-tet integration): algorithmic bytes / CUDA-event duration vs the measured
             HBM copy peak in MEASURED_PEAKS.json.
  cpu_baseline  the reference's CPU path on this box's host cores, on a bounded sample of the same workload.
  cpu_baseline_port  the flat restatement of that path (oracle/flat_oracle.cpp) on all host threads, same kind of sample.

--impl reference times the reference's own CPU implementation (oracle/_ref/dsc_ref, the unmodified
reference sources built by oracle/Makefile; falls back to the flat oracle port when that binary is absent)
on the same workload family, bounded so that the run ends within minutes.
Multi-GPU: one process per GPU (torchrun); tets / faces / nodes are partitioned, the volume and the snapshot
are replicated, and the only collectives are two NCCL all-reduces (per-phase sums, vertex forces).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "DSC image-force evals/sec (tets+faces/s)"
UNIT = "tets+faces/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="C4")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mean-mode", default="tet", choices=["tet", "zray"])
    ap.add_argument("--force-mode", default="auto", choices=["auto", "gather", "atomic"],
                    help="auto: the bit-exact ordered gather on one GPU, the face-partitioned atomic scatter when the forces are all-reduced anyway")
    ap.add_argument("--jitter", type=float, default=0.2)
    ap.add_argument("--dsc", default=None, help="replay a mesh saved by a reference run (iter_N.dsc, is_mesh/mesh_io.cpp:298-314) on the volume of --config instead of the generated grid mesh")
    ap.add_argument("--cpu-baseline-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap-geometry", action="store_true", help="one GPU: keep the geometry half of the force pass behind the tet pass (A/B)")
    ap.add_argument("--no-parity", action="store_true", help="skip the comparison of the step's outputs with the flat oracle (outside the timed region)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a captured CUDA graph")
    ap.add_argument("--overlap-forces", action="store_true",
                    help="start the mean-independent half of the force pass on a second stream next to the tet pass (measured: no gain, the tet kernel fills every SM)")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: exchange kernels over NVLink peer memory (csrc/dsc_comm.cuh) or two NCCL all-reduces")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ----------------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------------
def build_workload(pkg, cfg_name, seg_factory, jitter, dsc_path=None):
    syn = pkg.synthetic
    cfg = syn.CONFIGS[cfg_name]
    vol = syn.make_volume(cfg_name)
    img, seg = seg_factory(vol, cfg["nb_phase"])
    edge = syn.edge_for_cells(cfg["n"], cfg["cells"])
    if dsc_path:  # a mesh the reference saved: positions, tets and labels as they are in the file
        pos, tets, labels = pkg.mesh_io.read_dsc(dsc_path)
        return vol, cfg, edge, pkg.mesh_io.snapshot_from_tets(pos, tets, labels), img, seg

    def tet_mean(pos, tets, base):
        seg.set_snapshot({"pos": pos, "tet_nodes": tets, "tet_label": base, "face_nodes": np.zeros((0, 3), np.uint32),
                          "face_tets": np.zeros((0, 2), np.uint32), "face_is_boundary": np.zeros(0, np.uint8)})
        return seg.tet_integrate(include_bound=True, update_means=False)["mean"]

    snap = syn.build_case(vol, edge, cfg["nb_phase"], cfg["thresholds"], tet_mean, jitter=jitter, seed=5)
    return vol, cfg, edge, snap, img, seg


def algorithmic_bytes(snap, n_samples, n_gauss, voxel_bytes, per_tet_outputs=False):
    """per tet 16 B ids + 4 B label + 96 B positions + n^3 voxel reads (+24 B outputs when written);
    per interface face 60 B ids/labels + 120 B positions + 8 voxel reads per Gauss point + 144 B force read-modify-write."""
    n_tets = int((snap["tet_label"] != 0).sum())
    n_bound = len(snap["tet_label"]) - n_tets
    tet_b = n_tets * (16 + 4 + 96 + (24 if per_tet_outputs else 0)) + n_bound * 4 + n_samples * voxel_bytes
    face_b = len(snap["face_nodes"]) * (60 + 120 + 144) + n_gauss * 8 * voxel_bytes
    return tet_b, face_b


def parity_check(pkg, seg, vol, snap, P, sums, mean, forces, n_samples, bit_exact_forces, per_tet=True):
    """Outside the timed region: what the timed step produced against the flat oracle (oracle/flat_oracle.cpp, bit-identical to the
    unmodified reference on the fixtures) on the same inputs.  Integer decisions bit for bit, sums to the stated tolerance."""
    from oracle import flat
    thr = flat.num_threads()
    ref = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], threads=thr, want_hash=False)
    lab = snap["tet_label"]
    tot, volu, mean_ref = flat.phase_reduce(lab, ref["total"], ref["vol"], P)
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300)))
    out = {"oracle": "oracle/flat_oracle.cpp, %d threads" % thr,
           "sample_count_equal": bool(int(n_samples) == int(ref["nsamp"][lab != 0].sum())),
           "phase_total_max_rel": rel(sums[:P], tot), "phase_volume_max_rel": rel(sums[2 * P:3 * P], volu), "mean_max_rel": rel(mean, mean_ref)}
    # per-tet outputs of the default kernel: every sample in the reference's voxel <=> bit-equal totals for float voxels
    out["tet_volume_bit_equal"] = out["tet_total_mismatches"] = None
    if per_tet:  # (one GPU: with a partition set the per-tet call only covers this rank's range)
        o = seg.tet_integrate(include_bound=True, update_means=False)
        out["tet_volume_bit_equal"] = bool(np.array_equal(o["vol"], ref["vol"]))
        if vol.dtype == np.float32:
            out["tet_total_mismatches"] = int(np.count_nonzero(o["total"] != ref["total"]))
        else:  # integer voxels: the integer sum is divided at a different point, one rounding apart
            out["tet_total_mismatches"] = int(np.count_nonzero(np.abs(o["total"] - ref["total"]) > 1e-13 * np.abs(ref["total"]) + 1e-14))
    F_ref = flat.face_forces(vol, snap, P, np.asarray(mean, np.float64))  # one thread: the reference's order of additions
    F = np.asarray(forces).reshape(-1, 3)
    out["forces_bit_equal"] = bool(np.array_equal(F, F_ref))
    out["forces_max_abs_diff"] = float(np.abs(F - F_ref).max())
    out["forces_max_abs"] = float(np.abs(F_ref).max())
    tol_ok = out["phase_total_max_rel"] < 1e-9 and out["phase_volume_max_rel"] < 1e-9 and out["forces_max_abs_diff"] <= 1e-6 * max(out["forces_max_abs"], 1.0)
    out["ok"] = bool(out["sample_count_equal"] and out["tet_volume_bit_equal"] is not False and not out["tet_total_mismatches"] and tol_ok and (out["forces_bit_equal"] or not bit_exact_forces))
    return out


def flush_l2(torch, buf):
    buf.add_(1)  # read+write of a buffer larger than the 126 MB L2


# ----------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ----------------------------------------------------------------------------------------------------
def reference_binary():
    p = os.path.join(ROOT, "oracle", "_ref", "dsc_ref")
    return p if os.path.exists(p) else None


def run_reference_sample(pkg, cfg_name, target_seconds, jitter, force_port=False):
    """Times the reference CPU path on a bounded sample of the workload: the same phantom family and the same mesh
    edge length, on a sub-volume sized so that one pass takes roughly `target_seconds`.  Returns dict."""
    syn = pkg.synthetic
    cfg = syn.CONFIGS[cfg_name]
    edge = syn.edge_for_cells(cfg["n"], cfg["cells"])
    binp = None if force_port else reference_binary()
    # the reference builds an is_mesh complex (std::map based, ~6 us/tet) and samples ~1 us/tet: keep ~150 K tets
    n_sub = min(cfg["n"], 192 if target_seconds >= 10 else 128)
    if cfg_name == "C2":
        vol = syn.sphere_phantom(n_sub, radius=40.0 * n_sub / 128)
    else:
        lv = {"C3": ((0.1, 0.5, 0.9), (0.40, 0.24), 0.05), "C4": ((0.1, 0.5, 0.9), (0.40, 0.24), 0.02),
              "C5": ((0.1, 0.4, 0.65, 0.9), (0.42, 0.30, 0.16), 0.03)}[cfg_name]
        vol = syn.shells_phantom(n_sub, levels=lv[0], radii=lv[1], sigma=lv[2], seed=1234, dtype=cfg["dtype"])
    dt = {np.dtype(np.uint8): "u8", np.dtype(np.uint16): "u16", np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[vol.dtype]
    sample = "%s phantom cropped to %d^3 %s, same edge length %.3f, jitter %.2f" % (cfg_name, n_sub, dt, edge, jitter)
    if binp:
        with tempfile.TemporaryDirectory() as tmp:
            vp = os.path.join(tmp, "vol.raw")
            vol.tofile(vp)
            cmd = [binp, "--mode", "time", "--volume", vp, "--vol-dtype", dt, "--dims", str(n_sub), str(n_sub), str(n_sub), "--edge", repr(edge),
                   "--nb-phase", str(cfg["nb_phase"]), "--alpha", "0.01", "--labels", "thres", "--thres", ",".join(str(t) for t in cfg["thresholds"]),
                   "--jitter", str(jitter), "--jitter-seed", "5", "--repeat", "2"]
            t0 = time.time()
            out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
            wall = time.time() - t0
        r = json.loads(out.strip().splitlines()[-1])
        units = r["n_tets"] + r["n_if"]
        t_eval = r["t_tet_pass_s"] + r["t_external_force_s"]  # per-tet integration + per-phase means + face forces
        return {"value": units / t_eval, "unit": UNIT, "cores": 1, "kind": "reference",
                "sample": sample + "; unmodified reference sources (oracle/_ref/dsc_ref), get_tetra_intensity pass + compute_external_force, best of 2",
                "detail": r, "wall_s": wall, "units": units, "t_eval_s": t_eval}
    # fallback: flat oracle port, all host threads
    from oracle import flat
    thr = flat.num_threads()
    snap = syn.build_case(vol, edge, cfg["nb_phase"], cfg["thresholds"],
                          lambda pos, tets, base: flat.tet_integrate(vol, tets, pos, want_hash=False, threads=thr)["mean"], jitter=jitter, seed=5)
    mask = (snap["tet_label"] != 0).astype(np.uint8)
    best = 1e30
    for _ in range(2):
        t0 = time.time()
        o = flat.tet_integrate(vol, snap["tet_nodes"], snap["pos"], tet_mask=mask, want_hash=False, threads=thr)
        tot, volu, mean = flat.phase_reduce(snap["tet_label"], o["total"], o["vol"], cfg["nb_phase"])
        flat.face_forces(vol, snap, cfg["nb_phase"], mean, threads=thr)
        best = min(best, time.time() - t0)
    units = len(snap["tet_nodes"]) + len(snap["face_nodes"])
    return {"value": units / best, "unit": UNIT, "cores": thr, "kind": "port", "sample": sample + "; flat oracle port (oracle/flat_oracle.cpp), OpenMP",
            "units": units, "t_eval_s": best}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from __graft_entry__ import load_package
    pkg = load_package()
    vals = []
    info = None
    t_all = time.time()
    for i in range(args.warmup + args.steps):
        info = run_reference_sample(pkg, args.config, 4.0, args.jitter)
        if i >= args.warmup:
            vals.append(info["t_eval_s"])
        if time.time() - t_all > 240:  # keep the whole run within a few minutes
            break
    if not vals:
        vals = [info["t_eval_s"]]
    t = float(np.mean(vals))
    value = info["units"] / t
    cfg = pkg.synthetic.CONFIGS[args.config]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s: %d^3 %s phantom, %d phases (bounded sample, see cpu_baseline.sample)" % (
                args.config, cfg["n"], np.dtype(cfg["dtype"]).name, cfg["nb_phase"])},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------
def main_b200(args):
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the image-energy kernels have no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]  # the image's default prints a version banner to stdout, where exactly one JSON line is expected
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def seg_factory(vol, nb_phase):
        img = pkg.Image3D(vol, device=local_rank)
        return img, pkg.SegmentFunction(img, nb_phase, 0.01)

    t0 = time.time()
    vol, cfg, edge, snap, img, seg = build_workload(pkg, args.config, seg_factory, args.jitter, args.dsc)
    P = cfg["nb_phase"]
    n_tets, n_faces, n_nodes = len(snap["tet_nodes"]), len(snap["face_nodes"]), len(snap["pos"])
    units = n_tets + n_faces
    build_s = time.time() - t0
    mean_mode = pkg.MEAN_TET if args.mean_mode == "tet" else pkg.MEAN_ZRAY
    if args.force_mode == "auto":
        args.force_mode = "gather" if world == 1 else "atomic"
    force_mode = pkg.FORCE_GATHER if args.force_mode == "gather" else pkg.FORCE_ATOMIC

    # all launches go to one explicit (non-default) torch stream so that torch.cuda.Event brackets them;
    # NCCL collectives issued under torch.cuda.stream(stream) are ordered on the same stream
    stream = torch.cuda.Stream(device=dev)
    assert stream.cuda_stream != 0
    torch.cuda.set_stream(stream)
    img.set_stream(stream.cuda_stream)
    if mean_mode == pkg.MEAN_ZRAY:
        img.build_sum_table()
    seg.set_snapshot(snap)
    # weak scaling is not what north_star asks for C4 (fixed 2M-tet mesh split over N GPUs): strong scaling
    rng = lambda n: (n * rank // world, n * (rank + 1) // world)
    tet_range = rng(n_tets)
    if world > 1:
        # tets are cut by work, not by count: label-0 tets (the boundary layer, 8 % here, clustered at the ends and sides of the
        # tet array) cost almost nothing and a tet's cost is its sample count n^3 (DEMO/image3d.cpp:355-357: n = floor(cbrt(volume)))
        vol_t = seg.tet_integrate(include_bound=True)["vol"]
        n_lvl = np.clip(np.floor(np.cbrt(np.maximum(vol_t, 0.0)) + 1e-9), 1, 9)
        cost = np.where(snap["tet_label"] != 0, n_lvl ** 3 + 20.0, 0.5)
        cuts = pkg.distributed.balanced_cuts(cost, world)
        tet_range = (cuts[rank], cuts[rank + 1])
        seg.set_partition(tet_range, rng(n_faces), rng(n_nodes))
    p2p = world > 1 and args.comm == "p2p" and mean_mode == pkg.MEAN_TET
    if p2p:
        def gather_handles(h):
            out = [None] * world
            dist.all_gather_object(out, h)
            return out
        pkg.distributed.connect_peers(seg, rank, world, n_nodes, gather_handles, snapshot_bytes=32 * n_nodes + 20 * n_tets + 21 * n_faces + 4096)
        dist.barrier()  # every window exists and is mapped before the first exchange kernel runs
    if world == 1 and force_mode == pkg.FORCE_GATHER and mean_mode == pkg.MEAN_TET and not args.no_overlap_geometry:
        seg.set_overlap_geometry(True)  # the mean-independent half of the force pass runs next to the tail of the tet pass
    d_sums = torch.zeros(3 * P, dtype=torch.float64, device=dev)
    d_mean = torch.zeros(P, dtype=torch.float64, device=dev)
    d_forces = torch.zeros(3 * n_nodes, dtype=torch.float64, device=dev)
    l2_buf = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2
    seg.enable_timers(True)

    launches_per_step = [0]

    # the force array is zero except at nodes of interface faces (~14 % of the nodes): all-reduce only those rows
    active_idx = torch.from_numpy(np.unique(snap["face_nodes"]).astype(np.int64)).to(dev) if world > 1 else None
    forces_rows = d_forces.view(-1, 3)

    def allreduce_forces():
        packed = forces_rows.index_select(0, active_idx)
        dist.all_reduce(packed)
        forces_rows.index_copy_(0, active_idx, packed)

    def step_device():
        n0 = seg.launch_count()
        if p2p:  # tet pass + fused sums exchange, force scatter + fused row exchange: no NCCL, no pack / unpack kernels
            seg.tet_phase_sums_allreduce_dev(d_sums.data_ptr(), d_mean.data_ptr())
            seg.face_forces_allreduce_dev(d_mean.data_ptr(), d_forces.data_ptr())
            launches_per_step[0] = seg.launch_count() - n0
            return
        if args.overlap_forces and world == 1 and force_mode == pkg.FORCE_GATHER:
            seg.face_geometry_async()  # normals, areas, trilinear samples of the Gauss points: next to the tet pass, on the second stream
        if mean_mode == pkg.MEAN_TET:
            seg.tet_phase_sums_dev(d_sums.data_ptr())
        else:
            seg.zray_phase_sums_dev(d_sums.data_ptr())
        if world > 1:
            dist.all_reduce(d_sums)
        seg.means_from_sums_dev(d_sums.data_ptr(), d_mean.data_ptr())
        seg.face_forces_dev(d_mean.data_ptr(), d_forces.data_ptr(), force_mode)
        if world > 1:
            allreduce_forces()
        launches_per_step[0] = seg.launch_count() - n0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing: K steps, each bracketed by events, L2 flushed in between ----------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # samples clocks / throttle reasons from the warm-up to the end of the end-to-end loop
    t_warm = time.time()
    for _ in range(max(3, args.warmup)):
        flush_l2(torch, l2_buf)
        step_device()
    # keep the device busy for ~0.7 s of extra untimed steps so that the 20 ms clock sampler sees it under load
    while rank == 0 and world == 1 and time.time() - t_warm < 0.7:
        flush_l2(torch, l2_buf)
        step_device()
        torch.cuda.synchronize(dev)
    barrier()
    # ---- per-kernel durations (library CUDA events on the launching stream), eager launches ----
    tet_ms, force_ms = [], []
    for i in range(min(args.steps, 10)):
        flush_l2(torch, l2_buf)
        step_device()
        torch.cuda.synchronize(dev)
        tet_ms.append(seg.timer_ms("tet" if mean_mode == pkg.MEAN_TET else "zray"))
        force_ms.append(seg.timer_ms("force"))
    # ---- the step as one CUDA graph (kernels + NCCL all-reduces), replayed for the timed region: the evaluation is a
    # dozen short launches, so eager launch gaps would otherwise be part of every step.  Falls back to eager launches
    # when capture is not possible (the z-ray pass reads a hit count back mid-pass).
    graph = None
    if not args.no_graph and mean_mode == pkg.MEAN_TET:
        try:
            seg.enable_timers(False)
            torch.cuda.synchronize(dev)
            if world == 1 or p2p:
                g1 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1, stream=stream):
                    step_device()
                graph = [g1]
            else:
                # the NCCL all-reduces stay eager (not captured); the kernels on either side of them are two graphs
                ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(ga, stream=stream):
                    seg.tet_phase_sums_dev(d_sums.data_ptr())
                with torch.cuda.graph(gb, stream=stream):
                    seg.means_from_sums_dev(d_sums.data_ptr(), d_mean.data_ptr())
                    seg.face_forces_dev(d_mean.data_ptr(), d_forces.data_ptr(), force_mode)
                graph = [ga, gb]
            torch.cuda.synchronize(dev)
        except Exception as e:  # noqa: BLE001
            sys.stderr.write("CUDA graph capture failed (%r); timing eager launches\n" % (e,))
            graph = None
            seg.enable_timers(True)

    def step_graph():
        if world == 1 or p2p:
            graph[0].replay()
        else:
            graph[0].replay()
            dist.all_reduce(d_sums)
            graph[1].replay()
            allreduce_forces()

    if graph is not None:
        for _ in range(3):
            step_graph()
        torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for i in range(args.steps):
        flush_l2(torch, l2_buf)
        ev[i][0].record(stream)
        if graph is not None:
            step_graph()
        else:
            step_device()
        ev[i][1].record(stream)
    barrier()
    if graph is not None:
        seg.enable_timers(True)
    if p2p:
        seg.comm_error()  # a timed-out exchange would have produced garbage: fail loudly
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    n_samples = seg.last_sample_count()
    n_gauss = seg.last_gauss_count()
    total_ms = float(step_ms.sum())
    t = torch.tensor([total_ms, float(np.mean(tet_ms))], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(n_samples), float(n_gauss)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    total_ms_max, tet_ms_max = t.tolist()
    n_samples_all, n_gauss_all = [int(x) for x in cnt.tolist()]
    ms_per_step = total_ms_max / args.steps
    value = units / (ms_per_step * 1e-3)
    # what the timed step left on the device, kept for the parity check below
    res_sums = d_sums.cpu().numpy().copy()
    res_mean = d_mean.cpu().numpy().copy()
    res_forces = d_forces.cpu().numpy().copy()

    # ---- the same step with the per-label second moment (north_star (a): "intensity sum, squared sum and volume"; the reference's
    # get_tetra_intensity pass, and the reference arm, do not compute it, so `value` is quoted without it and this next to it) ----
    with_sumsq = None
    if mean_mode == pkg.MEAN_TET and (world == 1 or p2p):
        try:
            seg.set_hot_want_sq(True)
            seg.enable_timers(False)
            for _ in range(3):
                step_device()
            torch.cuda.synchronize(dev)
            g2 = None
            if graph is not None:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, stream=stream):
                    step_device()
                for _ in range(3):
                    g2.replay()
            ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            barrier()
            for i in range(args.steps):
                flush_l2(torch, l2_buf)
                ev2[i][0].record(stream)
                if g2 is not None:
                    g2.replay()
                else:
                    step_device()
                ev2[i][1].record(stream)
            barrier()
            t2 = torch.tensor([float(sum(a.elapsed_time(b) for a, b in ev2))], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t2, op=dist.ReduceOp.MAX)
            ms2 = t2.item() / args.steps
            with_sumsq = {"ms_per_step": ms2, "value": units / (ms2 * 1e-3), "unit": UNIT, "phase_sumsq": [float(x) for x in d_sums.cpu().numpy()[P:2 * P]]}
        finally:
            seg.set_hot_want_sq(False)
            seg.enable_timers(True)

    # ---- end to end through the C ABI with host buffers ------------------------------------------------------
    # Every step: the host flattens the mesh into the library's pinned staging arrays (here: a copy from the resident
    # numpy snapshot, standing in for the is_mesh walk), H2D, kernels, D2H of the statistics and the forces.
    st = seg.staging(n_nodes, n_tets, n_faces)
    pinned_forces = torch.empty(3 * n_nodes, dtype=torch.float64, pin_memory=True)
    pinned_mean = torch.empty(P, dtype=torch.float64, pin_memory=True)

    def fill_staging():
        st["pos4"][:, :3] = snap["pos"]
        st["tet_nodes"][...] = snap["tet_nodes"]
        st["tet_label"][...] = snap["tet_label"]
        st["face_nodes"][...] = snap["face_nodes"]
        st["face_tets"][...] = snap["face_tets"]
        st["face_is_boundary"][...] = snap["face_is_boundary"]

    t_fill = time.perf_counter()
    fill_staging()
    host_fill_ms = (time.perf_counter() - t_fill) * 1e3  # numpy copies into the pinned arrays; the adapter's is_mesh walk is timed in oracle/dropin_test.cpp
    h2d_bytes = n_nodes * 32 + n_tets * 20 + n_faces * (12 + 8 + 1)  # pos4, tet ids+labels, face ids/tets/flags; the CSR is built on the device
    n_active_nodes = len(np.unique(snap["face_nodes"]))
    # the forces come back as the rows of the nodes that have interface faces (+ their keys), every other row is zero; N > 1: on rank 0
    # (N = 1: one copy -- a header of 4 x 8 doubles for sums and means, the rows, the keys; N > 1: rows and keys on rank 0, the means on every rank)
    d2h_bytes = n_active_nodes * 28 + (4 * 8 * 8 if world == 1 else P * 8)

    def step_e2e():
        if world == 1:
            seg.staging(n_nodes, n_tets, n_faces)  # same pinned arrays (already filled); marks the snapshot dirty
            seg.image_force_eval(None, mean_mode=mean_mode, force_mode=force_mode, want_forces="compact")
        else:
            seg.staging(n_nodes, n_tets, n_faces)
            if p2p:  # every rank sends 1/world of the snapshot over PCIe, the shares cross NVLink (dscgpu_commit_snapshot_sharded)
                seg.commit_snapshot_sharded()
            else:
                seg.commit_snapshot()
            seg.set_partition(tet_range, rng(n_faces), rng(n_nodes))
            step_device()
            # every rank ends up with the same forces (rank-ordered sums): the job's result crosses PCIe once, on rank 0;
            # the other ranks read the means back (their synchronisation point)
            pinned_mean.copy_(d_mean, non_blocking=True)
            if rank == 0:
                seg.read_forces_compact(d_forces.data_ptr())  # rows of the nodes with interface faces + their keys (synchronises)
            torch.cuda.synchronize(dev)

    for _ in range(3):
        step_e2e()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = te.item() / e2e_steps * 1e3
    e2e_value = units / (e2e_ms * 1e-3)
    # where an end-to-end step spends its time: the library's own CUDA-event timers on one extra (untimed) step
    e2e_parts = None
    if world == 1:
        seg.enable_timers(True)
        t1 = time.perf_counter()
        step_e2e()
        wall = (time.perf_counter() - t1) * 1e3
        e2e_parts = {k + "_ms": round(seg.timer_ms(k), 4) for k in ("h2d", "tet", "phase_final", "force", "d2h")}
        e2e_parts["wall_ms_with_timers"] = round(wall, 4)
        e2e_parts["what"] = "h2d: first to last byte on the copy stream; tet: first staged launch to the last (they wait for their chunks); one extra step with event timers on"
        seg.enable_timers(False)
    # ---- the same with an incremental snapshot: every node position re-sent, 2 % of the tets replaced, the
    # interface-face list re-sent; reported next to e2e, never instead of it
    e2e_incr = None
    if world == 1:
        idx = np.arange(0, n_tets, 50, dtype=np.uint32)
        # like the full snapshot above: the host has flattened the update straight into the library's pinned arrays
        ust = seg.update_staging(n_nodes, False, len(idx), n_faces)
        ust["pos"][...] = snap["pos"]
        ust["tet_index"][...] = idx
        ust["tet_nodes"][...] = snap["tet_nodes"][idx]
        ust["tet_label"][...] = snap["tet_label"][idx]
        for k in ("face_nodes", "face_tets", "face_is_boundary"):
            ust[k][...] = snap[k]

        def step_incr():
            seg.update_commit(True)
            seg.image_force_eval(None, mean_mode=mean_mode, force_mode=force_mode, want_forces="compact")

        for _ in range(3):
            step_incr()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_incr()
        torch.cuda.synchronize(dev)
        incr_ms = (time.perf_counter() - t0) / e2e_steps * 1e3
        e2e_incr = {"value": units / (incr_ms * 1e-3), "unit": UNIT, "ms_per_step": incr_ms,
                    "h2d_bytes_per_step": int(n_nodes * 24 + len(idx) * 24 + n_faces * 21), "d2h_bytes_per_step": d2h_bytes,
                    "what": "dscgpu_update_commit from pinned staging: all node positions, 2 % of the tets, the interface-face list; then the evaluation on the resident snapshot"}
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel --------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    tet_bytes, face_bytes = algorithmic_bytes(snap, n_samples_all, n_gauss_all, vol.dtype.itemsize)
    tet_bytes_rank = tet_bytes / world
    kernel_ms = tet_ms_max
    achieved = tet_bytes_rank / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "tet_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            traffic = json.load(open(tp)).get(args.config)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": ("tet_staged_kernel + tet_exact_list_kernel" if seg.get_tet_variant() & 131072 else "tet_integrate_w2_kernel") if mean_mode == pkg.MEAN_TET else "zray pass", "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": tet_bytes_rank, "peak_source": peak_src,
                "whole_step": {"algorithmic_bytes": tet_bytes + face_bytes, "achieved": (tet_bytes + face_bytes) / world / (ms_per_step * 1e-3) / 1e9}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu_baseline = run_reference_sample(pkg, args.config, args.cpu_baseline_seconds, args.jitter)
            cpu_baseline = {k: cpu_baseline[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # the baseline is a reported extra; never fail the bench for it
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}
    # next to the reference on one core, the flat restatement on all host threads (a reported figure, never the target)
    cpu_port = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            cpu_port = run_reference_sample(pkg, args.config, 10.0, args.jitter, force_port=True)
            cpu_port = {k: cpu_port[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:  # noqa: BLE001 -- never fail the bench line over the second baseline
            cpu_port = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % (e,)}

    # ---- the estimator segment() really uses for the means (update_average_intensity: interface faces rasterised into z-rays over
    # the prefix-sum table) and the energy read-out, device-timed by the library's events; reported next to the metric, not in it ----
    zray = None
    if world == 1 and mean_mode == pkg.MEAN_TET:
        try:
            seg.set_snapshot(snap)
            img.build_sum_table()
            seg.enable_timers(True)
            m_z = seg.update_average_intensity()  # sizes the hit buffer
            means_ms, energy_ms = [], []
            for _ in range(5):
                flush_l2(torch, l2_buf)
                m_z = seg.update_average_intensity()
                means_ms.append(seg.timer_ms("zray"))
                flush_l2(torch, l2_buf)
                seg.compute_energy(mean=m_z)
                energy_ms.append(seg.timer_ms("zray"))
            zray = {"means_ms": float(np.median(means_ms)), "energy_ms": float(np.median(energy_ms)), "mean": [float(x) for x in m_z],
                    "phase_volume": [float(x) for x in seg._phase_volume], "what": "dscgpu_zray_means / dscgpu_zray_energy on the same snapshot, L2 flushed"}
        except Exception as e:  # noqa: BLE001
            zray = {"error": repr(e)}
    parity = None
    if not args.no_parity and mean_mode == pkg.MEAN_TET:
        try:
            parity = parity_check(pkg, seg, vol, snap, P, res_sums, res_mean, res_forces, n_samples_all, bit_exact_forces=(world == 1 and force_mode == pkg.FORCE_GATHER), per_tet=(world == 1))
        except Exception as e:  # noqa: BLE001 -- report, never hide
            parity = {"ok": False, "error": repr(e)}
    check = {"mean": [float(x) for x in d_mean.cpu().tolist()], "force_l2": float(torch.linalg.vector_norm(d_forces).item()),
             "force_sum": [float(x) for x in forces_rows.sum(dim=0).cpu().tolist()]}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %d^3 %s %d-phase phantom, %d tets (%d labelled), %d interface faces, %d nodes, %s" % (
            args.config, cfg["n"], vol.dtype.name, P, n_tets, int((snap["tet_label"] != 0).sum()), n_faces, n_nodes,
            ("mesh replayed from " + os.path.basename(args.dsc)) if args.dsc else "grid mesh edge %.3f jitter %.2f" % (edge, args.jitter)),
            "mean_mode": args.mean_mode, "force_mode": args.force_mode, "launch": "cuda graph replay" if graph is not None else "eager", "l2": "flushed between timed steps (256 MB write)",
            "partition": "tets (cut by sample count) / faces / nodes split in %d contiguous ranges, volume+snapshot replicated" % world,
            "comm": ("none" if world == 1 else ("fused exchange kernels over NVLink peer memory (CUDA IPC windows)" if p2p else "2 NCCL all-reduces")),
            "samples_per_step": n_samples_all, "gauss_points_per_step": n_gauss_all},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "ms_per_step": e2e_ms,
                "steps": e2e_steps, "host_fill_ms": host_fill_ms, "parts": e2e_parts},
        "with_sumsq": with_sumsq, "parity": parity, "zray": zray,
        "e2e_incremental": e2e_incr,
        "gpu_launches": int(launches_per_step[0] * args.steps),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "cpu_baseline_port": cpu_port, "clocks": clocks,
        "kernel_ms": {"tet_or_zray": float(np.mean(tet_ms)), "force": float(np.mean(force_ms)), "step_min": float(step_ms.min()),
                      "step_max": float(step_ms.max())},
        "setup_s": build_s, "check": check,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_b200(args)


if __name__ == "__main__":
    sys.exit(main())
