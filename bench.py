#!/usr/bin/env python
"""Benchmark of the sampling hot path: sampled ligands/sec, 1000-step DDPM.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one complete pass of the hot path over one batch: the 1000 reverse-diffusion steps
of KeypointDiffusion.sample_from_encoded_receptors for B complexes per GPU (graph build +
denoiser + posterior step per reverse step), from already-encoded pockets to coordinates + atom
features.  Default workload = BASELINE.json configs[1]: trained_models/gvp_20kp hyper-parameters,
1 synthetic pocket (20 keypoints), 100 ligands x 20 atoms, seeded random weights in the
reference state_dict layout (checkpoints and datasets are not available offline).

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline     dominant kernel (the fused edge kernel): algorithmic FLOPs / CUDA-event time,
               measured in an instrumented (non-graph) full trajectory right after the timed region
  cpu_baseline the CPU oracle (a port of the reference algorithm; the reference itself needs DGL /
               torch_cluster, which are not installed) on a bounded sample of the same workload
  e2e          same metric through the drop-in public API with HOST buffers (pinned), H2D + D2H inside
"""
import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DTYPES = {"bf16x3": "bf16x3 (split bf16 operand pairs on tcgen05, fp32 accumulate: fp32-grade, parity <= 1e-4)",
          "bf16": "bf16 operands on tcgen05, fp32 accumulate (~2e-3 of fp32; not the parity mode)",
          "fp32": "fp32 (SIMT FMA, the reference's own arithmetic)"}

WORKLOADS = {
    # name: (shipped config, pocket kind, n_kp, ligands per GPU, atoms per ligand)
    "gvp_20kp": ("gvp_20kp", "keypoint", 20, 100, 20),
    "egnn_20kp": ("egnn_20kp", "keypoint", 20, 100, 20),
    "egnn_20kp_c1": ("egnn_20kp", "keypoint", 20, 10, 20),       # BASELINE configs[0] (the CPU-runnable case)
    "egnn_40kp": ("egnn_40kp", "keypoint", 40, 800, 20),         # configs[2]: 6400 complexes over 8 GPUs
    "egnn_all_atom": ("egnn_all_atom", "all_atom", 500, 100, 20),  # configs[3]
    "gvp_ca": ("gvp_ca", "ca", 42, 1024, 20),                    # configs[4] (one point of the sweep)
}

# algorithmic FLOPs (DESIGN.md "Roofline")
def gvp_edge_flops(S=256, V=16, rbf=16, n_msg=3):
    f = 0
    for i in range(n_msg):
        vin = V + 1 if i == 0 else V
        h = max(vin, V)
        fin = S + rbf if i == 0 else S
        f += 2 * 3 * vin * h + 2 * 3 * h * V + 2 * (fin + h) * S + 2 * S * V
    return f


def egnn_edge_flops(H=257):
    # as the reference executes it: two branches of Linear(2H+1,H)+Linear(H,H), attention + coord heads
    return 2 * ((2 * H + 1) * H + H * H) * 2 + 2 * H + 2 * H


def egnn_edge_flops_min(H=257):
    # with the first Linear factorised onto the nodes (what the kernel does per edge)
    return 2 * 2 * H * H + 4 * H


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML, 200 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
                 "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def load_config(name):
    import yaml
    cfgs = yaml.safe_load(open(ROOT / "tests" / "golden" / "shipped_configs.yml"))
    return cfgs[name]


def make_pocket(kind, pocket_id, cfg, arch):
    from keypoint_diffusion_b200 import synthetic
    cut = cfg["graph"]["graph_cutoffs"]
    vs = cfg["dynamics_gvp"]["vector_size"] if arch == "gvp" else 0
    if kind == "keypoint":
        width = cfg["rec_encoder_gvp"]["out_scalar_size"] if arch == "gvp" else cfg["rec_encoder"]["out_n_node_feat"]
        return synthetic.keypoint_pocket(pocket_id, cfg["graph"]["n_keypoints"], width, vs, cut["kk"])
    if kind == "all_atom":
        return synthetic.all_atom_pocket(pocket_id, 500, len(cfg["dataset"]["rec_elements"]), vs, cut["rr"])
    return synthetic.ca_pocket(pocket_id, 42, len(cfg["dataset"]["rec_elements"]), vs, cut["rr"])


def build_model(cfg, device):
    from keypoint_diffusion_b200 import model_from_config
    os.chdir(ROOT)                       # dataset.location in the configs is relative
    torch.manual_seed(0)
    model = model_from_config(cfg)
    return model.to(device).eval() if device is not None else model.eval()


def oracle_step_time(model, cfg, arch, pocket, n_lig, n_timed, threads):
    """Mean seconds per reverse step of the CPU oracle on the same batch / weights.

    Bounded sample: n_timed reverse steps spread uniformly over s = 999..0 are executed and timed in
    full (graph build + denoiser + posterior step); the steps in between are fast-forwarded with the
    posterior update at eps = 0 (not timed), which keeps the state distribution of the trajectory --
    in particular the ligand-ligand edge count, which falls from fully connected at s = 999 as the
    untrained model lets the ligand expand -- the same as on the GPU arm."""
    from oracle import flat, schedule as OS
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import oracle_cfg
    torch.set_num_threads(threads)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    B = len(n_lig)
    if arch == "egnn":
        d = cfg["dynamics"]
        kw = dict(n_layers=d["n_layers"], hidden_nf=d["hidden_nf"], use_tanh=d["use_tanh"], message_norm=d["message_norm"],
                  update_kp_feat=d["update_kp_feat"], norm=d["norm"], ll_k=d["ll_k"], kl_k=d["kl_k"],
                  graph_cutoffs=cfg["graph"]["graph_cutoffs"])
        fwd = flat.egnn_forward
    else:
        d = cfg["dynamics_gvp"]
        kw = dict(vector_size=d["vector_size"], n_convs=d["n_convs"], n_hidden_scalars=d["n_hidden_scalars"],
                  message_norm=d["message_norm"], update_kp=d["update_kp"], ll_k=d["ll_k"], kl_k=d["kl_k"],
                  n_message_gvps=d["n_message_gvps"], n_update_gvps=d["n_update_gvps"], n_noise_gvps=d["n_noise_gvps"],
                  graph_cutoffs=cfg["graph"]["graph_cutoffs"])
        fwd = flat.gvp_forward
    ocfg = oracle_cfg(arch, kw, model.n_lig_features, model.n_kp_feat)
    nk = pocket.n_kp
    off = torch.arange(B).repeat_interleave(pocket.kk_src.numel()) * nk
    fb = flat.FlatBatch(lig_n=torch.tensor(n_lig), kp_n=torch.tensor([nk] * B), kp_x=pocket.kp_x.repeat(B, 1),
                        kp_h=pocket.kp_h.repeat(B, 1), kk_src=pocket.kk_src.repeat(B) + off,
                        kk_dst=pocket.kk_dst.repeat(B) + off,
                        kp_v=pocket.kp_v.repeat(B, 1, 1) if pocket.kp_v is not None else None)
    T = model.n_timesteps
    F = model.n_lig_features
    gamma = OS.gamma_table(T, 1e-5)
    g = torch.Generator().manual_seed(0)
    N_l = sum(n_lig)
    lig_b, kp_b = fb.batch_idx()
    fb.lig_x = torch.randn(N_l, 3, generator=g)
    fb.lig_h = torch.randn(N_l, F, generator=g)
    fb = flat.remove_com(fb, lig_b, kp_b, "ligand")
    timed = set(int(round(i * (T - 1) / max(n_timed - 1, 1))) for i in range(n_timed)) if n_timed > 1 else {T - 1}
    zeros = (torch.zeros(N_l, F), torch.zeros(N_l, 3))
    total = 0.0
    with torch.no_grad():
        for s_int in reversed(range(T)):
            nx, nh = torch.randn(N_l, 3, generator=g), torch.randn(N_l, F, generator=g)
            if s_int in timed:
                t0 = time.perf_counter()
                fb = flat.sample_p_zs_given_zt(lambda b, t: fwd(sd, ocfg, b, t), gamma, T, s_int, fb, nx, nh)
                total += time.perf_counter() - t0
            else:
                fb = flat.sample_p_zs_given_zt(lambda b, t: zeros, gamma, T, s_int, fb, nx, nh)
    return total / len(timed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="gvp_20kp", choices=sorted(WORKLOADS))
    ap.add_argument("--ligands", type=int, default=None, help="ligands per GPU (default: the workload's)")
    ap.add_argument("--steps-per-graph", type=int, default=50)
    ap.add_argument("--sub-batches", type=int, default=None,
                    help="concurrently sampled groups of complexes per GPU (default: the library's choice)")
    ap.add_argument("--cpu-steps", type=int, default=None, help="reverse steps per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16", "bf16x3"],
                    help="headline mode: bf16x3 (tcgen05 tensor cores, split bf16 operands, fp32-grade: inside the 1e-4 "
                         "parity bar), fp32 (SIMT, the reference's own arithmetic) or bf16 (tcgen05, plain bf16 operands, "
                         "GVP only, ~2e-3)")
    ap.add_argument("--no-mode-blocks", action="store_true",
                    help="skip the separately-reported measurements of the other precision modes")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name, pocket_kind, n_kp, B, n_atoms = WORKLOADS[args.workload]
    if args.ligands:
        B = args.ligands
    cfg = load_config(cfg_name)
    arch = cfg["diffusion"].get("architecture", "egnn")
    metric = "sampled ligands/sec (1000-step DDPM, 20 keypoints)"
    config = {"workload": f"trained_models/{cfg_name}: 1 synthetic pocket per GPU ({n_kp} keypoints), {B} ligands x "
                          f"{n_atoms} atoms per GPU, 1000 denoising steps, seeded random weights",
              "ligands_per_gpu": B, "atoms_per_ligand": n_atoms, "n_keypoints": n_kp, "timesteps": 1000,
              "parallelism": f"pocket/ligand sharding x{world}, one final gather",
              "l2": "state + weights (<60 MB) are L2-resident by design; every reverse step rewrites them"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        model = build_model(cfg, None)
        pocket = make_pocket(pocket_kind, 0, cfg, arch)
        n_lig = [n_atoms] * B
        n = args.cpu_steps or 12
        for _ in range(max(args.warmup, 0) and 1):
            oracle_step_time(model, cfg, arch, pocket, n_lig, 1, threads)
        times = [oracle_step_time(model, cfg, arch, pocket, n_lig, n, threads) for _ in range(args.steps)]
        t_step = sum(times) / len(times)
        value = B / (t_step * model.n_timesteps)
        sample = (f"per timed step: {n} reverse steps spread uniformly over s=999..0 executed in full (the others "
                  f"fast-forwarded with eps=0, untimed), batch {B}, mean step time x1000; CPU oracle "
                  f"(oracle/flat.py, a port: the reference needs DGL/torch_cluster), fp32, {threads} threads")
        print(json.dumps({"impl": "reference", "metric": metric, "value": value, "unit": "ligands/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * model.n_timesteps * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": value, "unit": "ligands/s", "cores": threads, "kind": "port",
                                           "sample": sample},
                          "e2e": {"value": value, "unit": "ligands/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from keypoint_diffusion_b200 import HeteroBatch, _lib, dist as kdist, ops
    import ctypes as C

    model = build_model(cfg, dev)
    pocket = make_pocket(pocket_kind, rank, cfg, arch)          # one pocket per rank
    n_lig = [n_atoms] * B
    g_host = HeteroBatch.from_pockets([pocket], n_lig, model.n_lig_features, pin=True)
    init_host = torch.zeros(B, 3).pin_memory()
    g_dev = g_host.to(dev)
    init_dev = init_host.to(dev)
    F = model.n_lig_features

    def one_sample_device():
        x, h = model.sample_from_encoded_receptors(g_dev, init_lig_pos=init_dev, seed=1234, sub_batches=args.sub_batches,
                                                   steps_per_graph=args.steps_per_graph, return_device_tensors=True)
        return kdist.gather_ligands(x, h, n_lig)

    def one_sample_e2e():
        g = HeteroBatch(g_host._bnn, g_host._ndata, g_host._edges, g_host._bne)    # fresh host view, same pinned tensors
        pos, feat = model.sample_from_encoded_receptors(g, init_lig_pos=init_host, seed=1234, sub_batches=args.sub_batches,
                                                        steps_per_graph=args.steps_per_graph)
        return pos, feat

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(k):
            fn()
        ev1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dt = max(ev0.elapsed_time(ev1) / 1e3, 0.0)
        barrier()
        # device time on the launching stream; the e2e leg ends with a blocking D2H, so wall >= device
        t = torch.tensor([dt, wall], device=dev, dtype=torch.float64)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    if args.precision != "fp32":
        model.dynamics.set_precision(args.precision)       # raises for a mode the architecture does not have
    launches0 = int(_lib.lib.kpd_launch_count())
    for _ in range(max(args.warmup, 0)):
        one_sample_device()
    clocks = ClockSampler(local_rank)
    clocks.start()
    dt, _ = timed(one_sample_device, args.steps)
    clk = clocks.finish()
    value = world * B * args.steps / dt
    lps = model.last_launches_per_step
    n_sub = args.sub_batches or model.default_sub_batches(B)
    config["sub_batches"] = (f"{n_sub} groups of complexes per GPU sampled concurrently (own CUDA graphs and streams; "
                             f"same noise as the undivided batch)")

    one_sample_e2e()
    dt_e, wall_e = timed(one_sample_e2e, args.steps)
    e2e_value = world * B * args.steps / max(dt_e, wall_e)
    kp = g_host.nodes["kp"].data
    h2d = sum(v.numel() * v.element_size() for v in kp.values()) + init_host.numel() * 4
    d2h = sum(n_lig) * (3 + F) * 4

    out = {"metric": metric, "value": value, "unit": "ligands/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": DTYPES[args.precision], "data": "synthetic", "config": config, "clocks": clk,
           "e2e": {"value": e2e_value, "unit": "ligands/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
           "reverse_steps_per_s": 1000 * args.steps / dt, "launches_per_reverse_step": lps}

    # ------------------------------------------------------------------ roofline leg (dominant kernel)
    if not args.no_roofline:
        prof_id = 2 if arch == "gvp" else 1
        n_layers = cfg["dynamics_gvp"]["n_convs"] if arch == "gvp" else cfg["dynamics"]["n_layers"]
        _lib.check(_lib.lib.kpd_profile_enable(prof_id, 1000 * n_layers + 16))
        torch.cuda.synchronize()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        model.sample_from_encoded_receptors(g_dev, init_lig_pos=init_dev, seed=1234, use_cuda_graph=False,
                                            return_device_tensors=True, sub_batches=1)
        pe1.record()
        torch.cuda.synchronize()
        serial_ms = pe0.elapsed_time(pe1)      # the instrumented trajectory: undivided batch, one stream, no CUDA graph
        tot, cnt = C.c_double(), C.c_int32()
        _lib.check(_lib.lib.kpd_profile_collect(C.byref(tot), C.byref(cnt)))
        _lib.lib.kpd_profile_enable(0, 0)
        prof_sampler = [s for s in model._samplers.values() if not isinstance(s, list)][-1]
        st = (C.c_double * 4)()
        _lib.check(_lib.lib.kpd_sampler_edge_stats(prof_sampler.handle, st))
        e_ll, e_kl, e_kk = st[0], st[1], st[2]
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md sustained)"
        if arch == "gvp":
            d = cfg["dynamics_gvp"]
            fe = gvp_edge_flops(d["n_hidden_scalars"], d["vector_size"], 16, d["n_message_gvps"])
            full = e_ll + 2 * e_kl + e_kk
            last = e_ll + e_kl
            flops_per_step = fe * (full * (n_layers - 1) + last) if d["update_kp"] else fe * last * n_layers
            kname = "gvp_edge_kernel" if args.precision == "fp32" else "gvp_edge_ws_kernel"
        else:
            H = cfg["dynamics"]["hidden_nf"] + 1
            fe = egnn_edge_flops_min(H)
            e_all = e_ll + (2 * e_kl + e_kk if cfg["dynamics"]["update_kp_feat"] else e_kl)
            flops_per_step = fe * e_all * n_layers
            kname = "egnn_edge_kernel" if args.precision == "fp32" else "egnn_edge_ws_kernel"
        n_rev = st[3] if st[3] > 0 else 1000.0
        avg_ms = tot.value / max(cnt.value, 1)
        achieved = (flops_per_step * n_rev / max(cnt.value, 1)) / (avg_ms * 1e-3) / 1e12 if avg_ms > 0 else 0.0
        traffic = None
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists():
            traffic = json.load(open(tj)).get(kname)
        out["roofline"] = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                           "frac": achieved / peak if peak else None, "traffic": traffic, "kernel": kname,
                           "peak_source": peak_src, "launches_timed": int(cnt.value), "avg_launch_ms": avg_ms,
                           "kernel_ms_per_reverse_step": tot.value / n_rev,
                           # share of the SERIAL instrumented trajectory (what an ncu launch list, which serialises
                           # launches, shows); in the timed region kernels of different sub-batches / edge types overlap
                           "kernel_share_of_step": tot.value / serial_ms,
                           "serial_ms_per_reverse_step": serial_ms / n_rev,
                           "flops_per_edge": fe, "mean_edges_per_step": {"ll": e_ll, "kl": e_kl, "lk": e_kl, "kk": e_kk},
                           "note": {"fp32": "fp32 SIMT tile GEMM; fraction is against the measured bf16 tensor peak",
                                    "bf16": "fused warp-specialised kernel: tcgen05 bf16 tile GEMMs + SIMT epilogues / gathers / "
                                            "segmented reduction; algorithmic FLOPs against the measured bf16 tensor peak",
                                    "bf16x3": "fused warp-specialised kernel; every algorithmic MAC costs FOUR bf16 tensor-core MACs "
                                              "(hi/lo operand pairs), so the tensor pipe does 4x the algorithmic FLOPs counted "
                                              "here; fraction is algorithmic FLOPs against the measured bf16 tensor peak"}[args.precision]}

    # ------------------------------------------------------------------ the other precision modes, stated separately
    if not args.no_mode_blocks:
        others = [p for p in (["bf16x3", "bf16", "fp32"] if arch == "gvp" else ["bf16x3", "fp32"]) if p != args.precision]
        acc = {"bf16x3": "denoiser output within ~3e-6 of the fp32 reference (bar 1e-4; tests/test_gpu_tensorcore.py)",
               "bf16": "denoiser output within ~2e-3 of fp32; not the parity mode",
               "fp32": "denoiser output within ~5e-7 of the fp32 reference (tests/test_gpu_parity.py)"}
        out["modes"] = {}
        for p in others:
            model.dynamics.set_precision(p)
            k = 1 if p == "fp32" else args.steps            # the SIMT mode is slow: one timed sample
            for _ in range(1 if p == "fp32" else 2):
                one_sample_device()
            dtp, _ = timed(one_sample_device, k)
            out["modes"][p] = {"value": world * B * k / dtp, "unit": "ligands/s", "ms_per_step": dtp / k * 1e3, "steps": k,
                               "dtype": DTYPES[p], "accuracy": acc[p]}
        model.dynamics.set_precision(args.precision)

    # gpu launches in the two timed regions: replayed graphs do not re-count, so derive from the captured sequence
    per_run = lps * 1000 + 9 * n_sub       # lps already sums the sub-batches' launches
    out["gpu_launches"] = int(per_run * args.steps * 2)          # headline mode: device-resident + e2e timed regions
    out["launch_counter_delta"] = int(_lib.lib.kpd_launch_count()) - launches0

    # ------------------------------------------------------------------ CPU baseline (rank 0, N=1)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        t1 = oracle_step_time(model, cfg, arch, pocket, n_lig, 1, threads)       # also warms the CPU code paths
        n = args.cpu_steps or max(2, min(40, int(20.0 / max(t1, 1e-3))))
        t_step = oracle_step_time(model, cfg, arch, pocket, n_lig, n, threads)
        out["cpu_baseline"] = {"value": B / (t_step * 1000), "unit": "ligands/s", "cores": threads, "kind": "port",
                               "sample": f"{n} reverse steps spread uniformly over s=999..0 of batch {B} executed in full "
                                         f"(others fast-forwarded with eps=0, untimed), mean step time x1000; "
                                         f"oracle/flat.py fp32, torch {torch.__version__}, {threads} threads"}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
