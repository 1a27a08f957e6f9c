"""Drop-in ``Renderer`` for the reference's plugin surface.

Select it from the reference's YAML (no reference file edited)::

    renderer_module: 'transhuman_b200.renderer'
    renderer_path: '<repo>/transhuman_b200/renderer.py'

``make_renderer`` (lib/networks/renderer/make_renderer.py:4-8) then calls
``Renderer(network)``; ``run.py:52,109`` call ``render_fast(batch)`` and
``if_nerf_clight.py:45`` calls ``render(batch)``.  Same names, arguments,
return dict (``rgb_map (1,N,3)``, ``acc_map (1,N)``, ``depth_map (1,N)``) and
batch keys as ``lib/networks/renderer/if_clight_renderer.py``.

Per frame, the encoder's ResNet backbone and the ViT stay the reference's torch modules (``net.encoder.model``,
``net.ViT``); everything else runs in ``libtranshuman_b200.so`` through :mod:`transhuman_b200.ops`: the encoder's tail
(upsample + cat + 1x1 convolutions, encoder.py:133-146) is evaluated inside the kernels that consume it -- SMPL
painting + cluster grouping from the latents (``th_paint_group_latents`` / ``th_group_mean``) and the
pre-mapped feature maps from the latents (``th_premap_from_latents``, 8f-2), so neither ``pixel_feat_map`` nor
``holder_feat_map`` is ever written -- then the whole per-sample-point path (``th_render_rays``).  An encoder that is
not the reference's ``SpatialEncoder`` in its default configuration is called as a black box and its maps go through
``th_paint_group`` / ``th_premap_features``.
Forward only: training (autograd + stratified jitter) keeps the reference path.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from . import ops

N_VERTS = 6890


class _DefaultCfg:
    """The hot-path knobs with the values of configs/train_or_eval.yaml:20-67."""
    N_samples = 64
    num_class = 300
    KNN = 7
    KNN_DIST_ALPHA = 0.5
    white_bkgd = False
    perturb = 0.
    rasterize = True
    time_steps = 1
    raw_noise_std = 0
    use_truncation = False
    embed_size = 192


def _resolve_cfg(cfg):
    if cfg is not None:
        return cfg
    try:  # running inside the reference tree as its plugin
        from lib.config import cfg as ref_cfg  # type: ignore
        return ref_cfg
    except Exception:
        return _DefaultCfg()


def segment_mean(x: torch.Tensor, pc2voxel: torch.Tensor, n_class: int) -> torch.Tensor:
    """Per-cluster mean over dim 0 -- the vectorised form of the reference's
    ``voxelization`` loop (if_clight_renderer.py:356-371; the dict keys are
    exactly ``arange(n_class)``).  Accumulated in float64."""
    flat = x.reshape(x.shape[0], -1).to(torch.float64)
    acc = torch.zeros((n_class, flat.shape[1]), dtype=torch.float64, device=x.device)
    acc.index_add_(0, pc2voxel, flat)
    cnt = torch.bincount(pc2voxel, minlength=n_class).to(torch.float64)
    return (acc / cnt[:, None]).reshape((n_class,) + tuple(x.shape[1:]))


class Renderer:
    """Same surface as ``if_clight_renderer.Renderer`` (39, 429, 486)."""

    def __init__(self, net, cfg=None, pc2voxel_ind=None, vertex_can=None):
        self.net = net
        self.cfg = _resolve_cfg(cfg)
        if vertex_can is None:
            # if_clight_renderer.py:43-48
            with open('./data/smplx/smpl/SMPL_NEUTRAL.pkl', 'rb') as f:
                vertex_can = pickle.load(f, encoding='latin1')['v_template']
        self.vertex_can = torch.as_tensor(np.asarray(vertex_can)).contiguous()
        self.CR = torch.tensor([-1.5, -1.5, -1.5, 1.5, 1.5, 1.5])
        num_voxel = int(self.cfg.num_class)
        dict_voxel2pc_ind = None
        if pc2voxel_ind is None:
            # if_clight_renderer.py:55
            d = np.load(f'./kmeans_dict/kmeans_dict_{num_voxel}.npy', allow_pickle=True).item()
            pc2voxel_ind = d['pc2voxel_ind']
            dict_voxel2pc_ind = d.get('dict_voxel2pc_ind')
        self.pc2voxel_ind = torch.as_tensor(np.asarray(pc2voxel_ind)).to(torch.int64)
        self.num_class = num_voxel
        assert int(self.pc2voxel_ind.max()) + 1 == num_voxel and len(torch.unique(self.pc2voxel_ind)) == num_voxel, \
            "cluster ids must be exactly arange(num_class)"
        # clusters in CSR form for the paint / grouping kernels (member order = the reference's dict lists)
        self.clusters = ops.ClusterIndex(pc2voxel_ind=self.pc2voxel_ind.numpy(), dict_voxel2pc_ind=dict_voxel2pc_ind,
                                         n_tok=num_voxel, device="cpu")
        # voxel_PE_can (if_clight_renderer.py:71): the literal per-cluster mean, once, on the CPU
        st, mem = self.clusters.start_host, torch.from_numpy(self.clusters.members_host.astype(np.int64))
        self.voxel_PE_can = torch.stack([self.vertex_can[mem[st[c]:st[c + 1]]].mean(0) for c in range(num_voxel)])
        self._weights = None
        self._weights_key = None
        self._mlp_params = None
        self.last_counters = None
        # 8f-2: run only the encoder's backbone and evaluate its tail inside the CUDA kernels (see prepare_frame);
        # False = call net.encoder as a black box (any encoder with the reference's four outputs)
        self.use_latents = True
        # 8f-3: the ViT's attention through th_vit_attention (flash-style); False = net.ViT as a black box
        self.use_flash_vit = True
        # replay the chains of small torch launches (encoder backbone, ViT blocks) as CUDA graphs (inference only)
        self.use_cuda_graphs = True
        self._graphs = {}
        # the ViT's Linear layers through th_linear from this many token rows (views x tokens) on
        self.use_tc_linear = True
        self.tc_linear_min_rows = 4096
        self._linears = {}
        # set `profile = True` to have every prologue stage bracketed by CUDA events; `last_prologue_ms` then
        # holds {stage: milliseconds} of the last prepare_frame (bench.py's `plugin` record)
        self.profile = False
        self.last_prologue_ms = {}
        self._stage_events = []

    # ---- prologue (torch; out of scope of the CUDA path) ---------------------------------
    def normalize_PE(self, PE):
        # if_clight_renderer.py:373-383
        lo, hi = self.CR[:3][None, None].to(PE.device), self.CR[3:][None, None].to(PE.device)
        return ((((PE - lo) / (hi - lo)) - 0.5) * 2).type(torch.float32)

    def _weights_fingerprint(self):
        """(data_ptr, in-place version) of every per-point-network parameter: changes when a checkpoint is loaded
        (load_state_dict copies in place -> version bump) or a parameter is replaced.  32 small tensors."""
        if self._mlp_params is None:
            self._mlp_params = [prm for name, prm in self.net.named_parameters()
                                if not name.startswith(('encoder.', 'ViT.'))]
        return tuple((prm.data_ptr(), prm._version) for prm in self._mlp_params)

    def _packed_weights(self, V, device):
        key = (V, str(device), self._weights_fingerprint() if hasattr(self.net, 'named_parameters') else None)
        if self._weights is None or self._weights_key != key:
            self._weights = ops.PackedWeights(self.net.state_dict(), V, device=device)
            self._weights_key = key
        return self._weights

    def _encoder_tail(self, images):
        """The reference SpatialEncoder's backbone (encoder.py:100-131) -> ops.EncoderTail, or None when
        ``net.encoder`` is not that module in its default configuration (then the caller runs it whole)."""
        enc = self.net.encoder
        m = getattr(enc, 'model', None)
        ok = (m is not None and all(hasattr(m, a) for a in ('conv1', 'bn1', 'relu', 'maxpool', 'layer1', 'layer2'))
              and getattr(enc, 'num_layers', None) == 3 and getattr(enc, 'feature_scale', None) == 1.0
              and getattr(enc, 'upsample_interp', None) == 'bilinear' and getattr(enc, 'index_interp', '') != 'nearest '
              and hasattr(enc, 'upsample_color') and hasattr(enc, 'reduction_layer')
              and tuple(enc.reduction_layer.weight.shape[:2]) == (192, 384))
        if not ok:
            return None
        def backbone(img):
            # channels_last memory format: cuDNN's native layout, and the latents come out channel-last in place -- the
            # layout th_premap_from_latents / th_paint_group_latents read (same values; only the strides differ)
            x = m.relu(m.bn1(m.conv1(img.contiguous(memory_format=torch.channels_last))))
            lat = [x]
            if enc.use_first_pool:
                x = m.maxpool(x)
            x = m.layer1(x)
            lat.append(x)
            lat.append(m.layer2(x))
            return tuple(lat)

        latents = self._graphed(("encoder", tuple(images.shape), str(images.device),
                                 tuple((p.data_ptr(), p.dtype) for p in m.parameters())), backbone, images)
        if [l.shape[1] for l in latents] != [64, 64, 128]:
            return None
        return ops.EncoderTail(list(latents), images, enc.upsample_color.weight, enc.upsample_color.bias)

    def _graphed(self, key, fn, *inputs):
        """``fn(*inputs)`` (a tuple of tensors, or one tensor) replayed as ONE CUDA graph per key -- the chains of
        small launches of the reference's torch modules are bound by Python + launch latency, not by the GPU.  Static
        input / output buffers; parameters are read through their own storage, so in-place weight updates are seen.
        Inference only (no grad); ``use_cuda_graphs = False`` runs eagerly."""
        if not self.use_cuda_graphs or torch.is_grad_enabled():
            return fn(*inputs)
        ent = self._graphs.get(key)
        if ent is None:
            static_in = [x.clone() for x in inputs]
            side = torch.cuda.Stream(device=inputs[0].device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):          # warm-up off the capture (cuDNN / cuBLAS workspaces, autotune)
                for _ in range(2):
                    fn(*static_in)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = fn(*static_in)
            ent = self._graphs[key] = (graph, static_in, static_out)
            if len(self._graphs) > 16:             # shapes rarely change; do not hoard graphs
                self._graphs.pop(next(iter(self._graphs)))
        graph, static_in, static_out = ent
        for dst, src in zip(static_in, inputs):
            dst.copy_(src)
        graph.replay()
        # the static outputs are overwritten by the next replay: hand out copies
        return tuple(o.clone() for o in static_out) if isinstance(static_out, tuple) else static_out.clone()

    def _vit_forward(self, tokens, pe):
        """``net.ViT(tokens, pe, mask=None)`` (vision_transformer.py:371-383) with the attention of every block
        through ``th_vit_attention`` (8f-3: no (V,3,N,N) attention matrix); LayerNorms, Linear layers and the MLP
        stay the reference's modules.  Anything that is not the reference's VisionTransformer at inference
        (dropout / stochastic depth active, other head sizes) is called whole."""
        vit = self.net.ViT
        blocks = getattr(vit, 'blocks', None)
        ok = self.use_flash_vit and blocks is not None and hasattr(vit, 'prepare_tokens') and hasattr(vit, 'norm')
        if ok:
            for blk in blocks:
                a = getattr(blk, 'attn', None)
                ok = ok and a is not None and all(hasattr(a, n) for n in ('qkv', 'proj', 'num_heads', 'scale')) \
                    and a.qkv.out_features == 3 * a.num_heads * 64 \
                    and all(hasattr(blk, n) for n in ('norm1', 'norm2', 'mlp')) \
                    and all(hasattr(blk.mlp, n) for n in ('fc1', 'fc2', 'act')) \
                    and (not vit.training or getattr(getattr(blk.mlp, 'drop', None), 'p', 0) == 0) \
                    and (not vit.training or (a.attn_drop.p == 0 and a.proj_drop.p == 0
                                              and isinstance(blk.drop_path, torch.nn.Identity)))
        if not ok:
            return vit(tokens, pe, mask=None)
        return self._graphed(("vit", tuple(tokens.shape), str(tokens.device),
                              tuple((p.data_ptr(), p.dtype) for p in vit.parameters())), self._vit_blocks, tokens, pe)

    def _vit_blocks(self, tokens, pe):
        vit = self.net.ViT
        x = vit.prepare_tokens(tokens, pe, None)
        # from a few thousand rows on, the four Linear layers of a block also go through the tensor cores
        # (th_linear: fp16 hi/lo three-product tcgen05 GEMM); below that cuBLAS' fp32 kernels are launch-bound anyway
        tc = self.use_tc_linear and x.shape[0] * x.shape[1] >= self.tc_linear_min_rows
        lin = self._packed_linear if tc else (lambda m: m)
        for blk in vit.blocks:
            a = blk.attn
            x = x + lin(a.proj)(ops.vit_attention(lin(a.qkv)(blk.norm1(x)), a.num_heads, a.scale))
            h = blk.mlp.act(lin(blk.mlp.fc1)(blk.norm2(x)))
            x = x + lin(blk.mlp.fc2)(h)
        return vit.norm(x)

    def _packed_linear(self, mod):
        """``ops.PackedLinear`` of an ``nn.Linear``, re-packed when its parameters change."""
        fp = tuple((p.data_ptr(), p._version) for p in (mod.weight, mod.bias) if p is not None)
        ent = self._linears.get(id(mod))
        if ent is None or ent[0] != fp:
            ent = self._linears[id(mod)] = (fp, ops.PackedLinear(mod.weight, mod.bias, device=mod.weight.device))
        return ent[1]

    def refresh_weights(self):
        """Forces a re-pack (only needed after REPLACING parameter tensors of ``net``; in-place updates such as
        ``load_state_dict`` are detected through the tensors' version counters)."""
        self._weights = None
        self._mlp_params = None

    def _stage(self, name):
        """Marks the start of a prologue stage (CUDA event on the current stream when profiling)."""
        if self.profile:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self._stage_events.append((name, e))

    def _stages_done(self):
        if not self.profile:
            return
        self._stage("end")
        torch.cuda.synchronize()
        ms = {}
        for (name, e0), (_, e1) in zip(self._stage_events[:-1], self._stage_events[1:]):
            ms[name] = ms.get(name, 0.0) + e0.elapsed_time(e1)
        self.last_prologue_ms = ms
        self._stage_events = []

    def prepare_frame(self, batch) -> ops.Frame:
        """encoder -> paint -> group -> ViT -> tokens (if_clight_renderer.py:531-547)."""
        assert int(getattr(self.cfg, 'time_steps', 1)) == 1
        dev = batch['ray_o'].device
        self._stage_events = []
        self._stage("pack_weights")
        images = batch['input_imgs'][0].reshape(-1, *batch['input_imgs'][0].shape[2:])
        weights = self._packed_weights(images.shape[0], dev)
        image_shape = batch['input_imgs'][0].shape[-2:]
        V = images.shape[0]
        viz = batch['input_vizmaps'][0][0] if (getattr(self.cfg, 'rasterize', True) and 'input_vizmaps' in batch) else None
        cams = (batch['input_R'][0].reshape(-1, 3, 3), batch['input_T'][0].reshape(-1, 3),
                batch['input_K'][0].reshape(-1, 3, 3))
        self._stage("encoder")
        tail = self._encoder_tail(images) if (self.use_latents and V <= 3) else None
        if tail is not None:
            # 8f-2: the encoder stops after its backbone; the upsample + cat + 1x1 convolutions of its tail
            # (encoder.py:133-146) are evaluated per vertex / per pixel inside th_paint_group_latents and the pre-map
            # GEMM, so neither pixel_feat_map (V,384,H,W) nor holder_feat_map (V,192,H,W) is written
            enc = self.net.encoder
            scale = np.array([images.shape[-1], images.shape[-2]], dtype=np.float64)
            holder_scale = pixel_scale = scale / (scale - 1) * 2.0   # encoder.py:149-153 (both maps are image-sized)
            hs = holder_scale / np.array(image_shape)
            self._stage("paint_group")
            grouped = ops.paint_group_latents(tail, enc.reduction_layer.weight, enc.reduction_layer.bias,
                                              (np.float32(hs[0]), np.float32(hs[1])),
                                              batch['input_smpl_vertice'][0][0], *cams, viz, self.clusters)
        else:
            holder_map, holder_scale, pixel_map, pixel_scale = self.net.encoder(images)
            # 8f-1: project + sample + visibility + cluster mean in ONE kernel (th_paint_group); token coordinates
            # and blend matrices through th_group_mean (bit-equal to the reference's voxelization on the CPU)
            self._stage("paint_group")
            hs = np.asarray(holder_scale, dtype=np.float64) / np.array(image_shape)
            grouped = ops.paint_group(holder_map, (np.float32(hs[0]), np.float32(hs[1])),
                                      batch['input_smpl_vertice'][0][0], *cams, viz, self.clusters)
        pe = self.voxel_PE_can.to(dev).unsqueeze(0).repeat(V, 1, 1)
        tok_xyz = ops.group_mean(batch['tar_smpl_vertice_smplcoord'][0].float().contiguous(), self.clusters)
        blend = batch['blend_mtx'][0]
        tok_rot = ops.group_mean(blend.contiguous(), self.clusters)[:, :3, :3].float() if blend.dtype == torch.float64 \
            else ops.group_mean(blend.float().contiguous(), self.clusters, outer_order=False)[:, :3, :3]
        self._stage("vit")
        holder = self._vit_forward(grouped.contiguous(), self.normalize_PE(pe))
        fs = np.asarray(pixel_scale, dtype=np.float64)
        sc = fs / np.array(image_shape)
        # pre-mapped maps (alpha_res_0 / rgb_res_0 / rgb_res_1 applied to the maps once per frame, tcgen05 GEMM over
        # the encoder's NCHW output) wherever the layer-chained schedule exists; plain channel-last maps otherwise
        self._stage("premap")
        premapped = V <= 3
        if tail is not None:
            feat = ops.premap_from_latents(tail, weights)
        else:
            feat = ops.premap_features(pixel_map, weights) if premapped else ops.nchw_to_nhwc(pixel_map)
        self._stages_done()
        return ops.Frame(
            holder=holder, tok_xyz=tok_xyz, tok_rot=tok_rot, verts=batch['tar_smpl_vertice'][0],
            feat_nhwc=feat, premapped=premapped, cam_R=batch['input_R'][0].reshape(-1, 3, 3),
            cam_T=batch['input_T'][0].reshape(-1, 3), cam_K=batch['input_K'][0].reshape(-1, 3, 3),
            Rh=batch['Rh'][0], Th=batch['Th'][0].reshape(3), weights=weights,
            uv_scale=(np.float32(sc[0]), np.float32(sc[1])), knn=int(self.cfg.KNN),
            knn_dist_alpha=float(self.cfg.KNN_DIST_ALPHA), white_bkgd=bool(self.cfg.white_bkgd))

    # ---- the reference's two entry points ------------------------------------------------
    def _check_forward_only(self):
        if float(getattr(self.cfg, 'perturb', 0.)) > 0. and getattr(self.net, 'training', False) \
                and torch.is_grad_enabled():
            raise NotImplementedError(
                "transhuman_b200.Renderer is forward-only: training (stratified jitter + autograd, "
                "if_clight_renderer.py:276-283) keeps the reference renderer")

    def _run(self, batch, mode):
        self._check_forward_only()
        with torch.no_grad():
            frame = self.prepare_frame(batch)
            out = ops.render_rays(frame, batch['ray_o'][0], batch['ray_d'][0], batch['near'][0], batch['far'][0],
                                  int(self.cfg.N_samples), mode=mode)
        self.last_counters = out["counters"]
        return {'rgb_map': out['rgb_map'][None], 'acc_map': out['acc_map'][None],
                'depth_map': out['depth_map'][None]}

    def render(self, batch, is_train=True):
        """if_clight_renderer.py:486-498 (every sample evaluated)."""
        return self._run(batch, ops.TH_RENDER_DENSE)

    def render_fast(self, batch, is_train=True):
        """if_clight_renderer.py:429-484.  Unlike the reference this does not
        overwrite batch['ray_o','ray_d','near','far'] with their compacted
        versions (459-462); nothing downstream reads them."""
        return self._run(batch, ops.TH_RENDER_FAST)
