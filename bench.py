"""bench.py -- molecules/s of the FlowMol sampling hot path on B200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload geom512|qm9_1024|dev_qm9_32]

One "step" = ONE PASS OF THE HOT PATH OVER ONE BATCH: a complete `integrate` of the batch (all `timesteps` Euler/CTMC
steps: T-1 network evaluations + the self-conditioning pre-pass + T-1 CTMC updates).  Default workload = BASELINE.json
configs[2], the configuration the metric is quoted on: GEOM-drugs sized molecules (n ~ train_data_n_atoms_histogram,
torch.Generator().manual_seed(1234)), flowmol3 dims, 512 molecules PER GPU, 250 timesteps, random-init weights
(seed 0), all-mask CTMC prior + COM-free N(0,1) positions.  N > 1: torchrun, one rank per GPU, each rank samples its share of (weak scaling): one global batch of 512 x N molecules is cut into contiguous cost-balanced ranges
(flowmol_b200/sharding.py), global molecule ids key the noise (=> same molecules at any N); the only collective is the final
NCCL gather of the results to rank 0, inside the timed region.

Printed JSON line (rank 0): metric/value/unit/..., `e2e` (same metric through fm_sample_host with pinned HOST buffers:
H2D of the prior + batch descriptor + trajectory + D2H of the result inside the timed region), `roofline` for the dominant
kernel family (the persistent tcgen05 linears k_egemm_p / _g / _c: timed live with CUDA events, per-mode numbers from in-pipeline events),
`cpu_baseline` (the CPU oracle port on a bounded sample),
`clocks`, `gpu_launches`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config, dataset, molecules per GPU, timesteps, BASELINE.json config index)
    "geom512": ("flowmol3", "geom", 512, 250, 2),
    "qm9_1024": ("flowmol3", "qm9", 1024, 250, 1),
    "dev_qm9_32": ("dev", "qm9", 32, 50, 0),
}
# BASELINE.json configs[4], the graph-irregularity stress: GEOM-type model, 2048 molecules IN TOTAL (strong scaling: 2048 / N per
# GPU), 250 timesteps, sizes either all equal to n ("sweep_n<n>") or uniform on {n/2 .. 3n/2} ("sweep_mix<n>"), n = 10 .. 80
SWEEP_TOTAL = 2048
for _n in (10, 20, 30, 40, 50, 60, 70, 80):
    WORKLOADS[f"sweep_n{_n}"] = ("flowmol3", f"fixed:{_n}", SWEEP_TOTAL, 250, 4)
    WORKLOADS[f"sweep_mix{_n}"] = ("flowmol3", f"uniform:{_n}", SWEEP_TOTAL, 250, 4)
# matmul FLOPs of the reference per directed edge of one GVPConv message phase (624 502, flowmol3)
CONV_EDGE_FLOP_PER_EDGE = {"flowmol3": 624502, "dev": 2 * (201 * 64 + 84 * 64 * 2 + 3 * 64 * 16) + 6 * (21 * 29 + 25 * 16 + 2 * (16 * 24 + 20 * 16))}
# matmul FLOPs of one full network evaluation, flowmol3: per edge / per node
FWD_FLOP = {"flowmol3": (4874436, 6499384), "dev": (308680, 324660)}


def draw_sizes(dataset, B, seed=1234, rank=0):
    """Sizes of a synthetic batch: 'geom' / 'qm9' = the training-set histograms shipped with the reference
    (data/*/train_data_n_atoms_histogram.pt), 'fixed:<n>' = all n atoms, 'uniform:<n>' = uniform on {n/2, .., 3n/2}."""
    gen = torch.Generator().manual_seed(seed + 7919 * rank)
    if dataset.startswith("fixed:"):
        return np.full(B, int(dataset.split(":")[1]), dtype=np.int64)
    if dataset.startswith("uniform:"):
        n = int(dataset.split(":")[1])
        return torch.randint(max(2, n // 2), n + n // 2 + 1, (B,), generator=gen).numpy().astype(np.int64)
    from flowmol_b200.api import n_atoms_histogram
    nmap, counts = n_atoms_histogram(dataset)
    return nmap[torch.multinomial(counts / counts.sum(), B, replacement=True, generator=gen)].numpy().astype(np.int64)


def atom_types_of(dataset):
    return 6 if dataset == "qm9" else 11          # QM9: C H N O F + fake; GEOM (and the sweep): 10 elements + fake


def make_prior(n_atoms, A, seed):
    """x_0 COM-free N(0,1) (flowmol/data_processing/priors.py:27-35), all-mask tokens (priors.py:101-107)."""
    gen = torch.Generator().manual_seed(seed)
    n = torch.from_numpy(n_atoms)
    N, U = int(n.sum()), int((n * (n - 1) // 2).sum())
    x0 = torch.randn(N, 3, generator=gen)
    nbi = torch.arange(len(n)).repeat_interleave(n)
    x0 = x0 - (torch.zeros(len(n), 3).index_add_(0, nbi, x0) / n[:, None].float())[nbi]
    return x0, torch.full((N,), A, dtype=torch.uint8), torch.full((N,), 6, dtype=torch.uint8), torch.full((U,), 4, dtype=torch.uint8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_port_throughput(cfg_name, dataset, timesteps, budget_s=20.0, rank_seed=0):
    """CPU baseline: the oracle port (oracle/flowmol_oracle.py == the reference's algorithm restated, pinned bit-exactly
    to the reference run verbatim) on all host cores, on a bounded sample of the same workload, extrapolated linearly in
    network evaluations to `timesteps` (state the extrapolation: value = molecules / (sec_per_eval * timesteps))."""
    from flowmol_b200 import weights as WT
    from flowmol_b200.config import ModelConfig
    from oracle import flowmol_oracle as O
    cores = os.cpu_count() or 1
    A = atom_types_of(dataset)
    cfg = ModelConfig.named(cfg_name, A)
    sd = WT.init_state_dict(cfg, 0)
    om = O.OracleModel(cfg, sd)
    nb = 8 if cfg_name == "flowmol3" else 32
    n_atoms = draw_sizes(dataset, nb, rank=rank_seed)
    bt = O.make_batch(n_atoms)
    x0, a0, c0, e0 = make_prior(n_atoms, A, 1)
    a0, c0, e0 = a0.long(), c0.long(), e0.long()
    with torch.no_grad():
        # the op mix is thousands of small ATen calls: more threads is not faster.  Calibrate on one evaluation each and keep
        # the best thread count (reported as `cores`).
        best = (None, 1e30)
        for nt in sorted({min(cores, n) for n in (8, 16, 32, 64, cores)}):
            torch.set_num_threads(nt)
            om.forward(bt, x0, a0, c0, e0, 0.0, None)
            t1 = time.perf_counter()
            d = om.forward(bt, x0, a0, c0, e0, 0.3, None)
            dt1 = time.perf_counter() - t1
            if dt1 < best[1]:
                best = (nt, dt1)
            if dt1 > 4 * best[1]:
                break
        cores = best[0]
        torch.set_num_threads(cores)
        d = om.forward(bt, x0, a0, c0, e0, 0.0, None)          # warm-up (also the pre-pass shape)
        evals, t0 = 0, time.perf_counter()
        while True:
            d = om.forward(bt, x0, a0, c0, e0, 0.3, d)
            evals += 1
            if time.perf_counter() - t0 > budget_s or evals >= 50:
                break
        sec_per_eval = (time.perf_counter() - t0) / evals
    value = nb / (sec_per_eval * timesteps)
    return {"value": value, "unit": "molecules/s", "cores": cores, "host_cpus": os.cpu_count(), "kind": "port",
            "sample": f"{nb} {dataset}-sized molecules (N={bt.N}, E={bt.E}), {evals} network evaluations timed "
                      f"({sec_per_eval:.3f} s each), extrapolated linearly to {timesteps} evaluations per molecule"}


def cpu_full_trajectory(cfg_name, dataset, B, timesteps, repeats=3):
    """the CPU path IN FULL on a configuration small enough for it (configs[0]: dev dims, 32 QM9-sized
    molecules, 50 timesteps): complete `integrate` (all network evaluations + CTMC steps) of the oracle port, best of `repeats`;
    where the reference tree is present (the build container, not the GPU box) the verbatim reference `CTMCVectorField.integrate`
    is timed once beside it."""
    from flowmol_b200 import weights as WT
    from flowmol_b200.config import ModelConfig
    from oracle import flowmol_oracle as O
    A = atom_types_of(dataset)
    cfg = ModelConfig.named(cfg_name, A)
    sd = WT.init_state_dict(cfg, 0)
    om = O.OracleModel(cfg, sd)
    n_atoms = draw_sizes(dataset, B)
    bt = O.make_batch(n_atoms)
    x0, a0, c0, e0 = make_prior(n_atoms, A, 100)
    a0, c0, e0 = a0.long(), c0.long(), e0.long()
    cores = min(os.cpu_count() or 1, 16)
    torch.set_num_threads(cores)
    best = 1e30
    with torch.no_grad():
        for r in range(repeats):
            t0 = time.perf_counter()
            O.integrate(om, bt, x0, a0, c0, e0, timesteps, seed=1000 + r)
            best = min(best, time.perf_counter() - t0)
    out = {"value": B / best, "unit": "molecules/s", "cores": cores, "host_cpus": os.cpu_count(), "kind": "port",
           "sample": f"the whole workload: {B} {dataset}-sized molecules (N={bt.N}, E={bt.E}), full {timesteps}-step integrate, "
                     f"best of {repeats} ({best:.2f} s)"}
    try:
        from oracle import ref_loader as RL
        if RL.available():
            R = RL.load()
            vf_cfg, sc_cfg = RL.read_vector_field_cfg(cfg_name)
            m = RL.build_reference_model(vf_cfg, sc_cfg, n_atom_types=A)
            m.load_state_dict(sd, strict=False)
            g, nbi, ebi, uem = RL.build_reference_graph([int(v) for v in n_atoms], generator=torch.Generator().manual_seed(4))
            Nn, Uu = g.num_nodes(), int(uem.sum())
            g.ndata['a_0'] = R.priors.ctmc_masked_prior(Nn, A)
            g.ndata['c_0'] = R.priors.ctmc_masked_prior(Nn, 6)
            ep = R.priors.ctmc_masked_prior(Uu, 4)
            e_0 = torch.zeros(uem.shape[0], 5)
            e_0[uem] = ep
            e_0[~uem] = ep
            g.edata['e_0'] = e_0
            with torch.no_grad():
                t0 = time.perf_counter()
                m.integrate(g, nbi, upper_edge_mask=uem, n_timesteps=timesteps, stochasticity=None, high_confidence_threshold=None)
                dt = time.perf_counter() - t0
            out["verbatim_reference"] = {"value": B / dt, "unit": "molecules/s", "seconds": dt,
                                         "what": "the reference's own CTMCVectorField.integrate (flowmol/models/ctmc_vector_field.py:145-285), "
                                                 "imported unmodified with the in-repo DGL / torch_scatter stand-ins, one run"}
    except Exception as exc:                      # noqa: BLE001 -- the verbatim arm is a bonus, never a reason to lose the line
        out["verbatim_reference"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    return out


def run_reference_arm(args, wl):
    """`--impl reference`: the reference's CPU path on this box's host cores.  The reference is pure Python on DGL / pytorch-scatter /
    Lightning / rdkit, none installable here, so the arm is the oracle PORT (oracle/flowmol_oracle.py, pinned bit-exactly to the
    reference run verbatim): `impl_detail` says so.  configs[0] runs in full; the flowmol3 workloads time a bounded sample and
    extrapolate linearly in network evaluations (stated in `cpu_baseline.sample`)."""
    cfg_name, dataset, B, T, _ = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for i in range(args.warmup + args.steps):
        if args.workload == "dev_qm9_32":
            r = cpu_full_trajectory(cfg_name, dataset, B, T, repeats=1)
        else:
            r = cpu_port_throughput(cfg_name, dataset, T, budget_s=max(3.0, 60.0 / max(1, args.warmup + args.steps)), rank_seed=0)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    cb = dict(vals[-1], value=v)
    line = {"impl": "reference", "impl_detail": "cpu_oracle_port (no GPU involved; n_gpus only mirrors the request)",
            "metric": f"molecules/sec @{T} steps ({dataset.upper()} batch)", "value": v,
            "unit": "molecules/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * B / v, "higher_is_better": True, "scaling": scaling_of(args.workload), "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(args.workload, wl, args.gpus),
            "cpu_baseline": cb, "e2e": {"value": v, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def scaling_of(workload):
    return "strong" if workload.startswith("sweep_") else "weak"


def config_dict(name, wl, gpus):
    cfg_name, dataset, B, T, idx = wl
    strong = scaling_of(name) == "strong"
    per_gpu, total = (B // gpus, B) if strong else (B, B * gpus)
    sizes = {"geom": "n_atoms ~ data/geom_full_kekulized/train_data_n_atoms_histogram", "qm9": "n_atoms ~ data/qm9/train_data_n_atoms_histogram"}.get(
        dataset, f"n_atoms {dataset.replace('fixed:', 'all = ').replace('uniform:', 'uniform on {n/2 .. 3n/2}, n = ')}")
    return {"workload": f"BASELINE.json configs[{idx}]: {dataset} sized molecules, {cfg_name} dims, {total} molecules in total "
                        f"({per_gpu} per GPU), {T} timesteps", "name": name, "molecules_per_gpu": per_gpu, "global_molecules": total,
            "timesteps": T,
            "sizes": sizes + ", one global draw torch.Generator().manual_seed(1234), contiguous cost-balanced ranges per rank",
            "weights": "random init (reference constructors' distributions), seed 0", "parallelism": f"molecule sharding x{gpus}",
            "l2": "inputs larger than L2 (edge hidden state alone is ~0.6 GB per evaluation at geom512); no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="geom512", choices=sorted(WORKLOADS))
    ap.add_argument("--eg-cluster", type=int, default=None, help="k_egemm_p cluster size (1, 2, 4): multicast weight stream")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE", help="fm_set_option before the run (A/B measurements), e.g. --opt pdl=0")
    ap.add_argument("--vec-impl", type=int, default=None, help="1 register-resident vector stages (default), 0 shared-memory tile kernels")
    ap.add_argument("--timesteps", type=int, default=None, help="override the workload's timesteps (experiments only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api-e2e", action="store_true", help="skip the extra pass through FlowMolB200.sample")
    ap.add_argument("--cuda-graph", type=int, default=0)
    ap.add_argument("--conv-impl", type=int, default=None, help="0 fp32 CUDA-core kernels, 2 wide tcgen05 pipeline (default where built)")
    ap.add_argument("--tc-prec", type=int, default=None, help="operand format of the tensor-core linears: 1 scaled fp16 hi/lo (default), 0 3xTF32")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.timesteps:
        wl[3] = args.timesteps
    wl = tuple(wl)
    cfg_name, dataset, B, T, _ = wl
    if args.impl == "reference":
        return run_reference_arm(args, wl)

    import torch.distributed as dist
    from flowmol_b200 import weights as WT
    from flowmol_b200.config import ModelConfig
    from flowmol_b200.vector_field import CTMCVectorFieldB200
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    A = atom_types_of(dataset)
    cfg = ModelConfig.named(cfg_name, A)
    vf = CTMCVectorFieldB200(cfg, WT.init_state_dict(cfg, 0), device=dev)
    strong = scaling_of(args.workload) == "strong"
    B_global = B if strong else B * world              # strong scaling (the size sweep): the total is fixed, ranks share it
    if args.conv_impl is not None:
        vf.set_option("conv_impl", args.conv_impl)
    if args.tc_prec is not None:
        vf.set_option("tc_prec", args.tc_prec)
    if args.eg_cluster is not None:
        vf.set_option("eg_cluster", args.eg_cluster)
    for kv in args.opt:
        vf.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    if args.vec_impl is not None:
        vf.set_option("vec_impl", args.vec_impl)
    # one GLOBAL batch of B x world molecules, cut into contiguous cost-balanced ranges (flowmol_b200/sharding.py); global
    # molecule ids key the noise, so the molecules are the same at any world size
    from flowmol_b200 import sharding as SH
    n_global = draw_sizes(dataset, B_global)
    ranges = SH.partition(n_global, world)
    lo, hi = ranges[rank]
    n_atoms = n_global[lo:hi]
    N, U, E = int(n_atoms.sum()), int((n_atoms * (n_atoms - 1) // 2).sum()), int((n_atoms * (n_atoms - 1)).sum())
    gx0, ga0, gc0, ge0 = make_prior(n_global, A, 100)
    noff = np.concatenate([[0], np.cumsum(n_global)])
    uoff = np.concatenate([[0], np.cumsum(n_global * (n_global - 1) // 2)])
    x0, a0, c0, e0 = gx0[noff[lo]:noff[hi]], ga0[noff[lo]:noff[hi]], gc0[noff[lo]:noff[hi]], ge0[uoff[lo]:uoff[hi]]
    hx, ha, hc, he = x0.clone().pin_memory(), a0.clone().pin_memory(), c0.clone().pin_memory(), e0.clone().pin_memory()
    dx, da, dc, de = x0.to(dev), a0.to(dev), c0.to(dev), e0.to(dev)

    def one_pass_device(seed):
        out = vf.integrate_tokens(n_atoms, dx, da, dc, de, T, seed=seed, mol_id_offset=lo, cuda_graph=bool(args.cuda_graph))
        if world > 1:                                     # the one exchange step: results to rank 0 over NCCL / NVLink
            SH.gather_results(out, n_global, ranges, rank, world)
        return out

    def one_pass_host(seed):
        bx, ba, bc, be = hx.clone().pin_memory(), ha.clone().pin_memory(), hc.clone().pin_memory(), he.clone().pin_memory()
        vf.sample_host(n_atoms, bx, ba, bc, be, T, seed=seed, mol_id_offset=lo, cuda_graph=bool(args.cuda_graph))
        return bx

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0_, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0_.record()
        for i in range(k):
            fn(1000 + i)
        e1_.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0_.elapsed_time(e1_)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0])

    for i in range(args.warmup):
        one_pass_device(i)
    with ClockSampler(local) as clk:
        ms_total = timed(one_pass_device, args.steps)
    launches = vf.last_launches * args.steps
    ms_per_step = ms_total / args.steps
    value = B_global / (ms_per_step / 1000.0)
    one_pass_host(0)
    ms_e2e = timed(one_pass_host, args.steps) / args.steps
    e2e_value = B_global / (ms_e2e / 1000.0)
    # the same pass through the reference-shaped public API (FlowMolB200.sample: graph construction, prior, integrate at the graph
    # seam, D2H, one SampledMolecule per molecule) by wall clock, one pass on rank 0's share (informational: `e2e` is the C-ABI call)
    e2e_api = None
    if not args.no_api_e2e:
        import flowmol_b200 as flowmol
        model = flowmol.FlowMolB200.from_config(cfg_name, dataset="qm9" if dataset == "qm9" else "geom", seed=0, device=str(dev),
                                                vector_field=vf)       # same weights (seed 0): adopt the built handle / workspace
        torch.manual_seed(0)
        nt = torch.from_numpy(n_atoms)
        model.sample(nt[:8], n_timesteps=4)           # warm the Python path
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        mols = model.sample(nt, n_timesteps=T)
        torch.cuda.synchronize()
        dt_api = time.perf_counter() - t0
        e2e_api = {"value": len(mols) * world / dt_api if not strong else len(mols) / dt_api, "unit": "molecules/s",
                   "seconds": dt_api, "what": "model.sample(n_atoms, n_timesteps) wall clock on rank 0's share, scaled by the rank count "
                                              "under weak scaling; includes graph build, prior sampling, D2H and molecule decode"}
    # roofline of the dominant kernel, timed live (CUDA events on its own stream) after a forward left valid state in the workspace
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    bf16 = peaks.get("bf16_tflops_sustained") or 1400.0
    peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md 1.4 PF sustained)"
    tf32_peak = bf16 / 2.0
    ffma_peak = 148 * 128 * 2 * (peaks.get("sm_max_mhz", 1965.0) * 1e6) / 1e12
    impl = vf.get_option("conv_impl")
    ms_pass = vf.time_conv_edge(layer=1, iters=3)
    prec = vf.get_option("tc_prec") if impl == 2 else 0
    if impl == 2:
        # Default flowmol3 pipeline.  The dominant kernel is k_egemm_p (persistent tcgen05 linear, 7 modes, ~53 % of a step); with
        # the fp16 (hi, lo) operand-image hand-over its modes are HBM-bound: roofline = algorithmic bytes of the mode (the fp32
        # rows the linear must read and write per edge) / live CUDA-event time /
        # measured HBM peak.  Headline entry: the 292 -> 256 message linear (EG_MSG; EG_MSGA is the same linear + segment-sum),
        # timed alone on its own stream; every mode's in-pipeline time comes from fm_debug_kprof events of one warm evaluation.
        ms_k = vf.time_egemm_msg(layer=1, iters=5)
        hbm_peak = peaks.get("hbm_gbs") or 6650.0
        hbm_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks.get("hbm_gbs") else "fallback (B200_PROFILING.md 6.65 TB/s)"
        prof = vf.kernel_profile(n_atoms, dx, da, dc, de)
        fwd_ms = sum(t_ for _, t_ in prof.values())
        # algorithmic HBM bytes per edge (fp32 rows read + written; per-node operands gathered from L2 are not counted) and matmul
        # FLOP per edge of every mode
        F_, S_ = 128, 256
        gate_fused = vf.get_option("eg_fuse_gate") == 1 and prof.get("EG_GATE", (0, 0))[0] <= 6.5
        modes = {"EG_MSG0": (4 * (F_ + 37) + 4 * S_, 2 * 197 * S_), "EG_GATE": (4 * S_ + 4 * 32, 2 * S_ * 32),
                 "EG_EU1": (4 * F_ + 4 * F_, 2 * 160 * F_), "EG_EU2": (4 * F_ + 4 * F_ + 4 * F_, 2 * F_ * F_),
                 # EdgeUpdate in one kernel (egemm_c.cuh): image in, fp32 residual in, fp32 rows + image out; both linears
                 "k_egemm_c": (4 * F_ * 4, 2 * 160 * F_ + 2 * F_ * F_)}
        if gate_fused:      # k_egemm_g: the gate linear rides along (+ 128 B of gate rows); GVP 2 writes no activation at all
            modes.update({"EG_MSG": (4 * 292 + 4 * S_ + 4 * 32, 2 * 292 * S_ + 2 * S_ * 32), "EG_MSGA": (4 * 292 + 4 * 32, 2 * 292 * S_ + 2 * S_ * 32)})
        else:
            modes.update({"EG_MSG": (4 * 292 + 4 * S_, 2 * 292 * S_), "EG_MSGA": (4 * 292 + 4 * S_, 2 * 292 * S_)})
        # the vector stages and their per-edge bytes (vec_reg.cuh: VU 384 B, SH 160 / 144 B, GT 128 B per edge)
        modes.update({"k_vecr_a": (384 + 160, 0), "k_vecr_b": (384 + 128 + 384 + 160, 0), "k_vecr_c": (384 + 128, 0),
                      "k_vec_a": (480 + 160, 0), "k_vec_b": (480 + 128 + 480 + 160, 0), "k_vec_c": (480 + 128, 0)})
        family = {}
        fam_bytes, fam_ms, moved, fam_launches, bind_num = 0.0, 0.0, 0.0, 0.0, 0.0
        for name_, (bpe, fpe) in modes.items():
            if name_ in prof:
                cnt, ms_f = prof[name_]
                us = 1e3 * ms_f / cnt
                x3_ceiling = (bf16 if prec == 1 else tf32_peak) / 3.0
                family[name_] = {"launches_per_eval": cnt, "us_per_launch": us, "algorithmic_bytes_per_edge": bpe,
                                 "achieved_gbs": bpe * E / (us * 1e-6) / 1e9, "frac_of_hbm_peak": bpe * E / (us * 1e-6) / 1e9 / hbm_peak,
                                 "algorithmic_tflops": fpe * E / (us * 1e-6) / 1e12,
                                 "frac_of_tensor_x3_ceiling": fpe * E / (us * 1e-6) / 1e12 / x3_ceiling}
                if name_.startswith("EG_") or name_ == "k_egemm_c":     # fraction of whichever roofline binds this kernel, time-weighted below
                    bind_num += max(family[name_]["frac_of_hbm_peak"], family[name_]["frac_of_tensor_x3_ceiling"]) * ms_f
                moved += bpe * E * cnt
                if name_.startswith("EG_") or name_ == "k_egemm_c":
                    fam_bytes += bpe * E * cnt
                    fam_ms += ms_f
                    fam_launches += cnt
        # Headline: the whole k_egemm_p family of one evaluation, time-weighted (every edge-row launch of the persistent tcgen05
        # linear: 51 % of an evaluation), not its best mode; the single-kernel EG_MSG number timed alone stays beside it.
        achieved = fam_bytes / (fam_ms * 1e-3) / 1e9
        hbm_bytes = E * (292 + 256) * 4
        flops = 2 * 292 * 256 * E
        op_peak = bf16 if prec == 1 else tf32_peak
        fe_, fn2 = FWD_FLOP[cfg_name]
        alg_eval_bytes = 6100.0 * E + 40000.0 * N                    # irreducible bytes of a fully fused evaluation
        alg_eval_flops = fe_ * E + fn2 * N
        roofline = {"kernel": "persistent tcgen05 linears (k_egemm_p / k_egemm_g / k_egemm_c), all edge-row launches of one network evaluation, "
                              f"time-weighted ({'fp16x3' if prec == 1 else '3xTF32'} operands, operand images between linears, gate linears and "
                              "EdgeUpdate fused in-kernel through tensor memory)",
                    "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_kind": hbm_src, "algorithmic_bytes_per_launch": fam_bytes / max(1.0, fam_launches),
                    "ms_per_launch": fam_ms / max(1.0, fam_launches),
                    "traffic": None,        # bench.py never runs under a profiler: dram__bytes of the committed captures are in profiles/
                    # the fused kernels are no longer HBM-bound: per kernel the larger of (HBM fraction, fraction of the fp16x3 tensor
                    # ceiling), weighted by time -- `frac` above stays the plain HBM fraction of the family
                    "binding_frac_time_weighted": bind_num / max(fam_ms, 1e-9),
                    "single_kernel": {"kernel": "k_egemm_p<EG_MSG> alone on its stream (292->256 message linear, all edges)",
                                      "ms_per_launch": ms_k, "algorithmic_bytes_per_launch": hbm_bytes,
                                      "achieved_gbs": hbm_bytes / (ms_k * 1e-3) / 1e9, "frac": hbm_bytes / (ms_k * 1e-3) / 1e9 / hbm_peak,
                                      "algorithmic_flops_per_launch": flops, "achieved_tflops": flops / (ms_k * 1e-3) / 1e12,
                                      "tensor_frac_of_x3_ceiling": flops / (ms_k * 1e-3) / 1e12 / (op_peak / 3.0)},
                    "whole_evaluation": {
                        "ms_event_to_event": fwd_ms,
                        "bytes_moved_by_the_listed_kernels": moved, "moved_gbs": moved / (fwd_ms * 1e-3) / 1e9,
                        "moved_frac_of_hbm_peak": moved / (fwd_ms * 1e-3) / 1e9 / hbm_peak,
                        "algorithmic_bytes_fused": alg_eval_bytes, "moved_over_algorithmic": moved / alg_eval_bytes,
                        "algorithmic_frac_of_hbm_peak": alg_eval_bytes / (fwd_ms * 1e-3) / 1e9 / hbm_peak,
                        "algorithmic_flops": alg_eval_flops, "algorithmic_tflops": alg_eval_flops / (fwd_ms * 1e-3) / 1e12,
                        "tensor_frac_of_x3_ceiling": alg_eval_flops / (fwd_ms * 1e-3) / 1e12 / (op_peak / 3.0),
                        "note": "SURVEY 8d: fused, the evaluation is compute-bound (about 800 FLOP per irreducible HBM byte); this build "
                                "keeps activations between linears in HBM, so its kernels are HBM-bound and `moved_over_algorithmic` "
                                "is the price of not fusing the GVP chain"},
                    "family": family,
                    "kernel_ms_per_eval": {k: round(t_, 4) for k, (c_, t_) in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                    "message_pass": {"what": "whole message phase of one conv layer (8 launches: 4 vector stages, MSG0, GATE, MSG + gate, MSGA + gate + segment-sum)",
                                     "ms": ms_pass, "algorithmic_tflops": CONV_EDGE_FLOP_PER_EDGE[cfg_name] * E / (ms_pass * 1e-3) / 1e12}}
    else:
        flops = CONV_EDGE_FLOP_PER_EDGE[cfg_name] * E
        achieved = flops / (ms_pass * 1e-3) / 1e12
        roofline = {"kernel": "k_conv_edge (fused gather + rbf + 3 message GVPs + segment-sum, one conv layer)", "bound": "tensor",
                    "achieved": achieved, "peak": tf32_peak, "unit": "TFLOP/s", "frac": achieved / tf32_peak,
                    "peak_kind": f"dense TF32 tensor = 1/2 x bf16, {peak_src}; this kernel runs on the fp32 CUDA-core path "
                                 f"(FFMA peak {ffma_peak:.1f} TFLOP/s => frac_of_ffma {achieved / ffma_peak:.3f})",
                    "algorithmic_flops_per_launch": flops, "ms_per_launch": ms_pass, "traffic": None,
                    "hbm_gbs_of_kernel": (E * cfg.n_hidden_edge_feats * 4) / (ms_pass * 1e-3) / 1e9}
    fe, fn_ = FWD_FLOP[cfg_name]
    total_flops = (fe * E + fn_ * N) * T * world
    line = {"metric": f"molecules/sec @{T} steps ({dataset.upper()} batch)", "value": value, "unit": "molecules/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling_of(args.workload),
            "vs_baseline": None,
            "dtype": ("f32 (tensor-core linears as error-compensated " + ("scaled fp16 hi/lo x3" if prec == 1 else "3xTF32") +
                      " on tcgen05, fp32 accumulate)") if impl == 2 else "f32",
            "data": "synthetic", "config": config_dict(args.workload, wl, world),
            "e2e": {"value": e2e_value, "unit": "molecules/s", "h2d_bytes_per_step": int(N * 12 + 2 * N + U + 4 * len(n_atoms)),
                    "d2h_bytes_per_step": int(N * 12 + 2 * N + U), "ms_per_step": ms_e2e},
            "e2e_api": e2e_api,
            "gpu_launches": int(launches), "roofline": roofline,
            "model_tflops": total_flops / (ms_per_step * 1e-3) / 1e12, "batch": {"N": N, "E": E, "U": U}}
    if rank == 0:
        line["clocks"] = clk.summary()
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = (cpu_full_trajectory(cfg_name, dataset, B, T) if args.workload == "dev_qm9_32"
                                    else cpu_port_throughput(cfg_name, dataset, T))
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
