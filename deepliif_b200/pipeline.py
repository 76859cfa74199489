"""Tile-batched inference driver: uint8 tiles in (host) -> five uint8 outputs + posneg mask out (host).

Replaces the reference's per-tile loop ``inference -> run_wrapper -> run_dask`` (deepliif/models/__init__.py:
464-579, 258-361), which runs batch = 1 and converts PIL <-> tensor <-> PIL around every generator call, with a
batched device pipeline: one H2D copy of the uint8 tiles, on-GPU ``transform`` (data/__init__.py:133-138), the
generators in micro-batches sized for L2 residency, on-GPU ``tensor2im`` quantisation (util/util.py:130-135),
seg aggregation + ``create_posneg_mask`` (models/__init__.py:338, postprocessing.py:163-190), one D2H copy of
uint8 results.  Two graph shapes:

  flat     : out_i = G_i(x), i = 1..5 (BASELINE.json configs 1-2: five ResNet-9 heads on the IHC tile; the
             fifth is the Seg head)
  cascade  : mods_i = G_i(x) (i = 1..4); seg = sum_k w_k * GS_k(x | mods_k)   (DeepLIIF_model.py:175-203)
"""
import torch

from . import ops


class TilePipeline:
    def __init__(self, gens, segs=None, seg_weights=None, micro_batch=8, thresh=120, n_streams=3):
        """gens: list of callables fp32 NCHW -> fp32 NCHW (modalities).  segs: None (flat: last of `gens` is the
        seg head) or list of len(gens)+1 seg generators (cascade)."""
        self.gens, self.segs = list(gens), (list(segs) if segs is not None else None)
        n_seg = len(self.segs) if self.segs is not None else 1
        self.seg_weights = list(seg_weights) if seg_weights is not None else [1.0 / n_seg] * n_seg
        self.micro_batch, self.thresh = micro_batch, thresh
        self.n_streams, self._stream_cache = n_streams, {}

    @torch.no_grad()
    def forward_device(self, x):
        """x: fp32 NCHW on device.  Returns (list of modality fp32 NCHW, seg fp32 NCHW, seg_u8 NHWC, mask).

        Independent (micro-batch, generator) chains are issued round-robin on `n_streams` CUDA streams: the
        memory-bound passes of one chain (normalise/split, statistics) then overlap the tensor-core-bound
        convolutions of another, and wave-quantisation tails are filled."""
        N, _, H, W = x.shape
        dev = x.device
        mb = self.micro_batch if self.micro_batch > 0 else N
        n_mod = len(self.gens) - (1 if self.segs is None else 0)
        mods_out = [torch.empty((N, 3, H, W), dtype=torch.float32, device=dev) for _ in range(n_mod)]
        seg_out = torch.empty((N, 3, H, W), dtype=torch.float32, device=dev)
        segu8_out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=dev)
        mask_out = torch.empty((N, H, W), dtype=torch.uint8, device=dev)
        keep_parts = getattr(self, "_keep_parts", None) is not None and self.segs is not None
        self._parts_out = ([torch.empty((N, 3, H, W), dtype=torch.float32, device=dev) for _ in self.segs]
                           if keep_parts else None)
        main = torch.cuda.current_stream()
        streams = self._streams(dev)
        for st in streams:
            st.wait_stream(main)
        k = 0
        for s in range(0, N, mb):
            xs = x[s:s + mb]
            if self.segs is None:
                for i, g in enumerate(self.gens):
                    st = streams[k % len(streams)]; k += 1
                    with torch.cuda.stream(st):
                        o = g(xs)
                        if i < n_mod:
                            mods_out[i][s:s + mb].copy_(o)
                        else:
                            seg, seg_u8, mask = ops.seg_finish([o], [1.0], self.thresh)
                            seg_out[s:s + mb].copy_(seg); segu8_out[s:s + mb].copy_(seg_u8); mask_out[s:s + mb].copy_(mask)
            else:
                # cascade: modality chain i feeds seg generator i+1; the base seg generator reads the tile itself
                parts = [None] * len(self.segs)
                used = []
                for i in range(len(self.segs)):
                    st = streams[k % len(streams)]; k += 1
                    used.append(st)
                    with torch.cuda.stream(st):
                        if i == 0:
                            parts[0] = self.segs[0](xs)
                        else:
                            m = self.gens[i - 1](xs)
                            mods_out[i - 1][s:s + mb].copy_(m)
                            parts[i] = self.segs[i](m)
                        if keep_parts:
                            self._parts_out[i][s:s + mb].copy_(parts[i])
                fin = used[0]
                for st in used[1:]:
                    fin.wait_stream(st)
                with torch.cuda.stream(fin):
                    seg, seg_u8, mask = ops.seg_finish(parts, self.seg_weights, self.thresh)
                    seg_out[s:s + mb].copy_(seg); segu8_out[s:s + mb].copy_(seg_u8); mask_out[s:s + mb].copy_(mask)
                for st in used[1:]:
                    st.wait_stream(fin)      # parts stay alive until seg_finish has consumed them
        for st in streams:
            main.wait_stream(st)
        return mods_out, seg_out, segu8_out, mask_out

    def _streams(self, dev):
        if self.n_streams <= 1:
            return [torch.cuda.current_stream()]
        key = str(dev)
        if key not in self._stream_cache:
            self._stream_cache[key] = [torch.cuda.Stream(device=dev) for _ in range(self.n_streams)]
        return self._stream_cache[key]

    @torch.no_grad()
    def infer_u8(self, tiles_u8_host, out_host=None, want_parts=False):
        """tiles_u8_host: pinned uint8 [N,H,W,3].  Returns dict of pinned uint8 host tensors:
        'mods' [M,N,H,W,3], 'seg' [N,H,W,3], 'mask' [N,H,W] (+ 'parts' [K,N,H,W,3], the per-modality seg outputs,
        when want_parts).  Copies are part of the call (end-to-end path); the D2H copies are asynchronous on the
        current stream — synchronise before reading."""
        dev = torch.device("cuda", torch.cuda.current_device())
        x_u8 = tiles_u8_host.to(dev, non_blocking=True)
        x = ops.u8_to_f32(x_u8)
        self._keep_parts = [] if want_parts else None
        mods, seg, seg_u8, mask = self.forward_device(x)
        mods_u8 = torch.stack([ops.f32_to_u8(m) for m in mods]) if mods else torch.empty((0,) + tuple(seg_u8.shape),
                                                                                          dtype=torch.uint8, device=dev)
        if out_host is None:
            out_host = {"mods": torch.empty(mods_u8.shape, dtype=torch.uint8, pin_memory=True),
                        "seg": torch.empty(seg_u8.shape, dtype=torch.uint8, pin_memory=True),
                        "mask": torch.empty(mask.shape, dtype=torch.uint8, pin_memory=True)}
        out_host["mods"].copy_(mods_u8, non_blocking=True)
        out_host["seg"].copy_(seg_u8, non_blocking=True)
        out_host["mask"].copy_(mask, non_blocking=True)
        if want_parts:
            parts_u8 = torch.stack([ops.f32_to_u8(p) for p in self._parts_out])
            out_host["parts"] = torch.empty(parts_u8.shape, dtype=torch.uint8, pin_memory=True)
            out_host["parts"].copy_(parts_u8, non_blocking=True)
        return out_host

    @torch.no_grad()
    def infer_mods_u8(self, tiles_u8_host):
        """Modalities only (mod_only / seg_gen=False): no seg generators are run."""
        dev = torch.device("cuda", torch.cuda.current_device())
        x = ops.u8_to_f32(tiles_u8_host.to(dev, non_blocking=True))
        mb = self.micro_batch if self.micro_batch > 0 else x.shape[0]
        outs = [torch.cat([g(x[s:s + mb]) for s in range(0, x.shape[0], mb)]) for g in self.gens]
        mods_u8 = torch.stack([ops.f32_to_u8(m) for m in outs])
        host = torch.empty(mods_u8.shape, dtype=torch.uint8, pin_memory=True)
        host.copy_(mods_u8, non_blocking=True)
        return {"mods": host}
