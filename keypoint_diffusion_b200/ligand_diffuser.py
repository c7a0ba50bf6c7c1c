"""Drop-in KeypointDiffusion for the sampling hot path.

Same constructor, state_dict layout and sampling API as the reference's
models/ligand_diffuser.py:24-538; the 1000-step loop of sample_from_encoded_receptors runs as a
replayed CUDA graph inside libkpdiff_b200.so (ops.Sampler).  Training (`forward`, losses) is
out of scope (sampling only) and raises NotImplementedError.
"""
import os
from math import ceil
from pathlib import Path
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import hetero, ops
from .dynamics import LigRecDynamics, LigRecDynamicsGVP
from .n_nodes_dist import LigandSizeDistribution
from .receptor_encoder import ReceptorEncoder, ReceptorEncoderGVP
from .schedule import PredefinedNoiseSchedule, alpha, coefficient_table, sigma, sigma_and_alpha_t_given_s
from .utils import copy_graph, get_batch_idxs, split_bounds

DEFAULT_SUB_BATCHES = {"gvp": 4, "egnn": 2}      # measured on B200: profiles/r01_sweep_sub_batches_*.txt


class FixedReceptorEncoder(nn.Module):
    """reference models/receptor_encoder_fixed.py:9-66: keypoints := receptor atoms, kk := rr."""

    def __init__(self, n_vec_feats):
        super().__init__()
        self.n_vec_feats = n_vec_feats

    def forward(self, g, batch_idxs=None):
        rec = g.nodes["rec"].data
        nd = {"kp": {"x_0": rec["x_0"], "h_0": rec["h_0"]},
              "lig": dict(g.nodes["lig"].data),
              "rec": {k: v[:0] for k, v in rec.items()}}
        if self.n_vec_feats is not None:
            nd["kp"]["v_0"] = torch.zeros((rec["x_0"].shape[0], self.n_vec_feats, 3), device=g.device)
        zeros = torch.zeros_like(g.batch_num_nodes("rec"))
        rr = g.edges(form="uv", etype="rr")
        return hetero.HeteroBatch({"kp": g.batch_num_nodes("rec"), "lig": g.batch_num_nodes("lig"), "rec": zeros},
                                  nd, {("kp", "kk", "kp"): rr}, {("kp", "kk", "kp"): g.batch_num_edges("rr")})


class KeypointDiffusion(nn.Module):

    def __init__(self, atom_nf, rec_nf, processed_dataset_dir: Path, n_timesteps: int = 1000, keypoint_centered=False,
                 architecture: str = 'egnn', rec_encoder_type: str = 'learned', graph_config={}, dynamics_config={},
                 rec_encoder_config={}, rec_encoder_loss_config={}, precision=1e-4, lig_feat_norm_constant=1,
                 rl_dist_threshold=0, use_fake_atoms=False):
        super().__init__()
        self.n_lig_features = atom_nf
        self.n_kp_feat = rec_nf
        self.n_timesteps = n_timesteps
        self.lig_feat_norm_constant = lig_feat_norm_constant
        self.use_fake_atoms = use_fake_atoms
        self.rec_encoder_type = rec_encoder_type
        if architecture not in ['egnn', 'gvp']:
            raise ValueError(f'Unsupported architecture: {architecture}')
        self.architecture = architecture
        if use_fake_atoms:
            # dead and broken in the reference (torch.cumsum without dim, max_fake_atom_frac 0 everywhere)
            raise NotImplementedError("use_fake_atoms is not supported (no shipped config enables it)")
        self.lig_size_dist = LigandSizeDistribution(processed_dataset_dir=Path(processed_dataset_dir))
        self.gamma = PredefinedNoiseSchedule(noise_schedule='polynomial_2', timesteps=n_timesteps, precision=precision)
        dynamics_config = dict(dynamics_config)
        if 'no_cg' in rec_encoder_config:
            dynamics_config['no_cg'] = rec_encoder_config['no_cg']
        dynamics_class = LigRecDynamics if architecture == 'egnn' else LigRecDynamicsGVP
        self.dynamics = dynamics_class(atom_nf, rec_nf, **graph_config, **dynamics_config)
        if rec_encoder_type not in ['learned', 'fixed']:
            raise ValueError(f'Receptor encoder type must be either "learned" or "fixed". Got {rec_encoder_type=} instead.')
        if rec_encoder_type == 'learned':
            encoder_class = ReceptorEncoder if architecture == 'egnn' else ReceptorEncoderGVP
            self.rec_encoder = encoder_class(**graph_config, **rec_encoder_config)
        else:
            self.rec_encoder = FixedReceptorEncoder(rec_encoder_config['vector_size'] if architecture == 'gvp' else None)
        object.__setattr__(self, "_samplers", {})
        object.__setattr__(self, "_coef", {})
        object.__setattr__(self, "last_launches_per_step", 0)

    # ------------------------------------------------------------------ training (out of scope)
    def forward(self, complex_graphs, interface_points=None):
        raise NotImplementedError("training is outside the sampling hot path this package implements")

    # ------------------------------------------------------------------ small helpers with reference names
    def normalize(self, g):
        g.nodes['lig'].data['h_0'] = g.nodes['lig'].data['h_0'] / self.lig_feat_norm_constant
        return g

    def unnormalize(self, g):
        g.nodes['lig'].data['h_0'] = g.nodes['lig'].data['h_0'] * self.lig_feat_norm_constant
        return g

    def sigma(self, gamma):
        return sigma(gamma)

    def alpha(self, gamma):
        return alpha(gamma)

    def sigma_and_alpha_t_given_s(self, gamma_t, gamma_s):
        return sigma_and_alpha_t_given_s(gamma_t, gamma_s)

    def _layout(self, g):
        return self.dynamics._layout(g)

    def remove_com(self, g, lig_batch_idx=None, kp_batch_idx=None, com: str = None):
        """reference :185-203 -- in place on g's lig / kp x_0, through kpd_remove_com."""
        if com is None:
            raise NotImplementedError('removing COM of receptor/ligand complex not implemented')
        if com not in ('ligand', 'receptor'):
            raise ValueError(f'invalid value for com: {com=}')
        batch, _ = self._layout(g)
        x_lig = g.nodes['lig'].data['x_0'].float().contiguous()
        x_kp = g.nodes['kp'].data['x_0'].float().contiguous()
        ops.remove_com(batch, x_lig, x_kp, com)
        g.nodes['lig'].data['x_0'] = x_lig
        g.nodes['kp'].data['x_0'] = x_kp
        return g

    def encode_receptors(self, g):
        return self.rec_encoder(g, get_batch_idxs(g))

    def coef_table(self, device) -> torch.Tensor:
        key = (str(device), self.gamma.gamma._version, self.gamma.gamma.data_ptr())
        if key not in self._coef:
            self._coef.clear()
            self._coef[key] = coefficient_table(self.gamma.gamma, self.n_timesteps).to(device)
        return self._coef[key]

    # ------------------------------------------------------------------ one reverse step (reference :497-538)
    @torch.no_grad()
    def sample_p_zs_given_zt(self, s: torch.Tensor, t: torch.Tensor, g, batch_idxs=None, noise=None, seed: int = 0):
        """Mutates lig x_0/h_0 and kp x_0 of g in place and returns g.  ``noise=(pos_noise, feat_noise)``
        injects the Gaussian draws (parity); otherwise Philox keyed by (seed, step, atom, channel)."""
        batch, _ = self._layout(g)
        eps_h, eps_x = self.dynamics(g, t, batch_idxs)
        s_int = int(torch.round(s.flatten()[0] * self.n_timesteps).item())
        step = torch.tensor([s_int], dtype=torch.int32, device=g.device)
        x_lig = g.nodes['lig'].data['x_0'].float().contiguous()
        h_lig = g.nodes['lig'].data['h_0'].float().contiguous()
        x_kp = g.nodes['kp'].data['x_0'].float().contiguous()
        nx, nh = (None, None) if noise is None else (noise[0].float().contiguous(), noise[1].float().contiguous())
        ops.ddpm_step(batch, x_lig, h_lig, x_kp, eps_x, eps_h, self.coef_table(g.device), step, nx, nh, seed)
        g.nodes['lig'].data['x_0'], g.nodes['lig'].data['h_0'], g.nodes['kp'].data['x_0'] = x_lig, h_lig, x_kp
        return g

    # ------------------------------------------------------------------ the loop (reference :342-469)
    def _sampler(self, g, steps_per_graph, use_cuda_graph) -> ops.Sampler:
        batch, kk = self._layout(g)
        model = self.dynamics.device_model(g.device)
        key = (id(batch), id(kk), id(model), steps_per_graph, use_cuda_graph, getattr(model, "precision", "fp32"))
        if key not in self._samplers:
            if len(self._samplers) > 3:
                self._samplers.clear()
            self._samplers[key] = ops.Sampler(model, batch, self.dynamics.graph_params(), kk, self.coef_table(g.device),
                                              self.n_timesteps, self.n_lig_features, steps_per_graph=steps_per_graph,
                                              use_cuda_graph=use_cuda_graph,
                                              lig_feat_norm_constant=float(self.lig_feat_norm_constant))
        return self._samplers[key]

    def _sub_samplers(self, g, n_sub, steps_per_graph, use_cuda_graph):
        """The batch cut into n_sub contiguous groups of complexes, each with its own captured loop (ops.Sampler) and its
        own stream.  Complexes are independent (no cross-complex term anywhere in the loop), so the groups' kernels may
        run concurrently: a fused edge kernel of one group is only a few waves of long-running CTAs, and the CTAs of the
        other groups fill the SMs its tail leaves idle.  Noise is keyed by the GLOBAL atom index, so the split does not
        change what is drawn."""
        batch, kk = self._layout(g)
        model = self.dynamics.device_model(g.device)
        key = ("sub", id(batch), id(kk), id(model), n_sub, steps_per_graph, use_cuda_graph, getattr(model, "precision", "fp32"))
        if key not in self._samplers:
            if len(self._samplers) > 3:
                self._samplers.clear()
            B = batch.B
            ks, kd = g.edges(form="uv", etype="kk")
            ks, kd = ks.cpu().long(), kd.cpu().long()
            kp_ptr = batch.kp_ptr.cpu().long()
            lig_ptr = batch.lig_ptr.cpu().long()
            bounds = split_bounds(B, n_sub)
            subs = []
            for a, b in zip(bounds[:-1], bounds[1:]):
                sb = ops.DeviceBatch(batch.lig_n[a:b], batch.kp_n[a:b], g.device)
                k0, k1, l0, l1 = int(kp_ptr[a]), int(kp_ptr[b]), int(lig_ptr[a]), int(lig_ptr[b])
                m = (kd >= k0) & (kd < k1)
                skk = ops.Csr.from_edges(ks[m] - k0, kd[m] - k0, sb.n_kp, g.device)
                smp = ops.Sampler(model, sb, self.dynamics.graph_params(), skk, self.coef_table(g.device), self.n_timesteps,
                                  self.n_lig_features, steps_per_graph=steps_per_graph, use_cuda_graph=use_cuda_graph,
                                  lig_feat_norm_constant=float(self.lig_feat_norm_constant), atom_offset=l0)
                subs.append((smp, (a, b), (k0, k1), (l0, l1), torch.cuda.Stream(device=g.device)))
            self._samplers[key] = subs
        return self._samplers[key]

    def default_sub_batches(self, n_complexes: int) -> int:
        """How many concurrently sampled groups a batch is cut into when the caller does not say (measured on B200,
        DESIGN.md section 4.4); KPD_SUB_BATCHES overrides."""
        env = os.environ.get("KPD_SUB_BATCHES")
        if env:
            return max(1, min(int(env), n_complexes))
        return max(1, min(DEFAULT_SUB_BATCHES[self.architecture], n_complexes // 16))

    @torch.no_grad()
    def sample_from_encoded_receptors(self, g, visualize=False, init_lig_pos: torch.Tensor = None, noise=None,
                                      seed: Optional[int] = None, steps_per_graph: int = 50, use_cuda_graph: bool = True,
                                      return_device_tensors: bool = False, sub_batches: Optional[int] = None):
        """Returns (lig_pos, lig_feat): one CPU tensor per complex, as the reference does.  ``g`` may
        live on the CPU (pinned or not): its keypoint tensors are uploaded here, which is the
        host->device boundary of the path (reference test.py:152-161)."""
        if visualize:
            raise NotImplementedError("visualize=True (a per-step CPU copy of the whole graph) is a debug feature "
                                      "outside the throughput path")
        dev = torch.device("cuda", torch.cuda.current_device()) if g.device.type != "cuda" else g.device
        if g.device.type != "cuda":
            g = g.to(dev)
        batch_size = g.batch_size
        if init_lig_pos is not None:
            assert init_lig_pos.shape == (batch_size, 3)
            init_pos = init_lig_pos.to(dev, torch.float32, non_blocking=True)
        else:
            if g.num_nodes('rec') == 0:
                raise ValueError("init_lig_pos is required when the graph has no 'rec' nodes (fixed encoder; "
                                 "the reference would take a mean over zero nodes here, SURVEY N7)")
            init_pos = hetero.readout_nodes(g, feat='x_0', op='mean', ntype='rec')
        kp = g.nodes['kp'].data
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())      # follows torch.manual_seed like the reference
        n_sub = self.default_sub_batches(batch_size) if sub_batches is None else max(1, min(int(sub_batches), batch_size))
        if noise is not None:
            n_sub = 1          # injected noise (parity runs) is laid out for the undivided batch
        if n_sub == 1:
            sampler = self._sampler(g, steps_per_graph, use_cuda_graph)
            x_lig, h_lig, x_kp = sampler.run(kp['x_0'], kp['h_0'], kp.get('v_0'), init_pos, noise=noise, seed=seed)
            object.__setattr__(self, "last_launches_per_step", sampler.launches_per_step)
        else:
            subs = self._sub_samplers(g, n_sub, steps_per_graph, use_cuda_graph)
            cur = torch.cuda.current_stream(dev)
            kx, kh, kv = kp['x_0'].float().contiguous(), kp['h_0'].float().contiguous(), kp.get('v_0')
            kv = kv.float().contiguous() if kv is not None else None
            init_pos = init_pos.float().contiguous()
            ready = torch.cuda.Event()
            ready.record(cur)
            parts = []
            for smp, (a, b), (k0, k1), _, st in subs:
                st.wait_event(ready)
                with torch.cuda.stream(st):
                    out = smp.run(kx[k0:k1], kh[k0:k1], kv[k0:k1] if kv is not None else None, init_pos[a:b], seed=seed)
                for t in out:
                    t.record_stream(cur)
                parts.append(out)
            for *_, st in subs:
                cur.wait_stream(st)
            x_lig = torch.cat([p[0] for p in parts])
            h_lig = torch.cat([p[1] for p in parts])
            x_kp = torch.cat([p[2] for p in parts])
            object.__setattr__(self, "last_launches_per_step", sum(smp.launches_per_step for smp, *_ in subs))
        g.nodes['lig'].data['x_0'], g.nodes['lig'].data['h_0'], g.nodes['kp'].data['x_0'] = x_lig, h_lig, x_kp
        if return_device_tensors:
            return x_lig, h_lig
        sizes = g.batch_num_nodes('lig').tolist()
        pos_cpu, feat_cpu = x_lig.cpu(), h_lig.cpu()                # device -> host boundary (reference :464)
        return list(torch.split(pos_cpu, sizes)), list(torch.split(feat_cpu, sizes))

    # ------------------------------------------------------------------ batch drivers (reference :270-340, :472-495)
    @torch.no_grad()
    def _sample(self, ref_graphs: List, n_lig_atoms: List[List[int]], rec_enc_batch_size: int = 32,
                diff_batch_size: int = 32, visualize=False, use_ref_lig_com: bool = False, encoded: bool = False,
                init_lig_pos: Optional[List[torch.Tensor]] = None):
        """ref_graphs: one single-complex graph per receptor.  ``encoded=True`` skips the receptor
        encoder (graphs already hold kp nodes + kk edges).  init_lig_pos: optional [3] tensor per receptor."""
        n_receptors = len(ref_graphs)
        if encoded:
            enc_graphs = ref_graphs
        else:
            # encode the pockets rec_enc_batch_size at a time (reference :277-287), on the device the weights live on
            dev = self.gamma.gamma.device
            enc_graphs = []
            for b in range(ceil(n_receptors / rec_enc_batch_size)):
                chunk = hetero.batch(ref_graphs[b * rec_enc_batch_size:(b + 1) * rec_enc_batch_size])
                enc_graphs.extend(hetero.unbatch(self.encode_receptors(chunk if chunk.device == dev else chunk.to(dev))))
        graphs, centers = [], []
        for rec_idx, ref_graph in enumerate(enc_graphs):
            sizes = n_lig_atoms[rec_idx]
            graphs.extend(copy_graph(ref_graph, n_copies=len(sizes), lig_atoms_per_copy=torch.tensor(sizes)))
            if init_lig_pos is not None:
                centers.extend([init_lig_pos[rec_idx].reshape(1, 3)] * len(sizes))
        n_complexes = len(graphs)
        lig_pos, lig_feat = [], []
        for b in range(ceil(n_complexes / diff_batch_size)):
            sl = slice(b * diff_batch_size, min((b + 1) * diff_batch_size, n_complexes))
            bg = hetero.batch(graphs[sl])
            if init_lig_pos is not None:
                center = torch.cat(centers[sl])
            elif use_ref_lig_com:
                # copy_graph zero-fills ligand data, so this is the origin -- as in the reference
                center = hetero.readout_nodes(bg, feat='x_0', op='mean', ntype='lig')
            else:
                center = None
            p, f = self.sample_from_encoded_receptors(bg, visualize=visualize, init_lig_pos=center)
            lig_pos.extend(p)
            lig_feat.extend(f)
        samples, end = [], 0
        for rec_idx in range(n_receptors):
            start, end = end, end + len(n_lig_atoms[rec_idx])
            samples.append({'positions': lig_pos[start:end], 'features': lig_feat[start:end]})
        return samples

    @torch.no_grad()
    def sample_given_pocket(self, rec_graph, n_lig_atoms: torch.Tensor, rec_enc_batch_size: int = 32,
                            diff_batch_size: int = 32, visualize=False, **kw):
        samples = self._sample([rec_graph], n_lig_atoms=[n_lig_atoms.tolist()], rec_enc_batch_size=rec_enc_batch_size,
                               diff_batch_size=diff_batch_size, visualize=visualize, **kw)
        return samples[0]['positions'], samples[0]['features']

    @torch.no_grad()
    def sample_random_sizes(self, ref_graphs: List, n_replicates: int = 10, rec_enc_batch_size: int = 32,
                            diff_batch_size: int = 32, **kw):
        n_nodes_rec = torch.tensor([max(g.num_nodes('rec'), g.num_nodes('kp') if self.rec_encoder_type == 'fixed' else 0)
                                    for g in ref_graphs])
        n_lig_atoms = self.lig_size_dist.sample(n_nodes_rec, n_replicates).tolist()
        return self._sample(ref_graphs=ref_graphs, n_lig_atoms=n_lig_atoms, rec_enc_batch_size=rec_enc_batch_size,
                            diff_batch_size=diff_batch_size, **kw)
