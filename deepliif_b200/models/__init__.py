"""Model factory and inference helpers with the reference's public names and argument meaning
(deepliif/models/__init__.py:101-660): create_model, init_nets, run_dask, run_wrapper, inference, infer_modalities.

What changes underneath: the reference runs one tile at a time (batch 1, PIL <-> tensor around every net, Dask
threads across nets, nets spread over GPUs).  Here all tiles of an image form one uint8 batch that goes through
``TilePipeline`` (H2D once, on-GPU transform, all generators on the sm_100a kernels in micro-batches, on-GPU
quantisation, D2H of uint8 results); several GPUs shard tiles, not networks (deepliif_b200/sharding.py).
``infer_modalities`` ends, as in the reference, with ``postprocess`` (cell-level scoring, SegOverlaid, SegRefined), whose
pixel/graph work runs on the GPU (deepliif_b200/postprocessing.py).
Out of scope: executing TorchScript graphs (their weights do load), TorchServe, Dask, WSI readers."""
import importlib
import os
from functools import lru_cache

import numpy as np
import torch
from PIL import Image

from .base_model import BaseModel
from .. import ops
from ..options import Options, print_options
from ..pipeline import TilePipeline
from ..util import TileGrid, disable_batchnorm_tracking_stats, image_variance_gray
from ..util.util import tensor_to_pil

Image.MAX_IMAGE_PIXELS = None
EMPTY_TILE_VARIANCE = 9          # is_empty(): tiles with gray-level variance below this skip the networks


def find_model_using_name(model_name):
    """deepliif_b200.models.<model_name>_model.<ModelName>Model (case-insensitive), a BaseModel subclass."""
    try:
        modellib = importlib.import_module(f"deepliif_b200.models.{model_name}_model")
    except ImportError:
        raise NotImplementedError(f"model [{model_name}] is outside the B200 hot-path scope (DeepLIIF only)")
    target = model_name.replace("_", "") + "model"
    for name, cls in vars(modellib).items():
        if name.lower() == target.lower() and isinstance(cls, type) and issubclass(cls, BaseModel):
            return cls
    raise NotImplementedError(f"In {model_name}_model.py, there should be a subclass of BaseModel named {target}")


def create_model(opt):
    instance = find_model_using_name(opt.model)(opt)
    print("model [%s] was created" % type(instance).__name__)
    return instance


def load_eager_models(opt, devices=None):
    model = create_model(opt)
    model.setup(opt)
    nets = {}
    for name in (devices.keys() if devices else model.model_names):
        net = getattr(model, "net" + name)
        if opt.phase != "train":
            net = disable_batchnorm_tracking_stats(net.eval())
        nets[name] = net.module if hasattr(net, "module") and not hasattr(net, "engine") else net
        if devices:
            nets[name].to(devices[name])
    return nets


@lru_cache
def get_opt(model_dir, mode="test"):
    return Options(path_file=os.path.join(model_dir, "train_opt.txt"), mode=mode)


@lru_cache
def init_nets(model_dir, eager_mode=True, opt=None, phase="test"):
    """dict[name -> callable net] for a model directory (train_opt.txt + latest_net_*.pth), built once.
    A directory written by the reference's `deepliif serialize` (G1.pt, ..., TorchScript) loads too: the weights are
    extracted from the archives into the native layout (the traced cuDNN graphs themselves are not executed), so
    `eager_mode` only selects which files are looked for first."""
    if opt is None:
        opt = get_opt(model_dir, mode=phase)
    if opt.model != "DeepLIIF":
        raise Exception(f"init_nets() not implemented for model {opt.model}")
    opt.phase = phase
    if not hasattr(opt, "epoch"):
        opt.epoch = "latest"
    return load_eager_models(opt)


def _names(opt):
    S, base = opt.mod_id_seg, int(opt.input_id)
    gens = [f"G{i + 1}" for i in range(opt.modalities_no)]
    segs = [f"G{S}{i + base}" for i in range(opt.modalities_no + 1)] if opt.seg_gen else []
    return gens, segs


def _seg_weights(opt, seg_weights):
    if seg_weights is not None:
        return list(seg_weights)
    return [1 / (opt.modalities_no + 1)] * (opt.modalities_no + 1)


_PIPE_CACHE = {}


def _pipeline(nets, gen_keys, seg_keys, weights, micro_batch, n_streams):
    """One TilePipeline per (networks, pruning, weights): its captured CUDA graphs are reused across images."""
    key = (tuple(id(nets[k]) if k else None for k in gen_keys), tuple(id(nets[k]) if k else None for k in seg_keys or ()),
           tuple(weights), micro_batch, n_streams)
    pipe = _PIPE_CACHE.get(key)
    if pipe is None:
        if len(_PIPE_CACHE) > 8:
            _PIPE_CACHE.clear()
        pipe = TilePipeline([nets[k] if k else None for k in gen_keys],
                            [nets[k] if k else None for k in seg_keys] if seg_keys else None, weights,
                            micro_batch=micro_batch, n_streams=n_streams, use_graph=os.getenv("DLB_NO_GRAPH", "") == "")
        _PIPE_CACHE[key] = pipe
    return pipe


_PINNED = {}


def _pinned(tag, shape):
    """Reusable pinned staging buffer (cudaHostAlloc costs milliseconds: never per image)."""
    n = int(np.prod(shape))
    buf = _PINNED.get(tag)
    if buf is None or buf.numel() < n:
        buf = _PINNED[tag] = torch.empty(max(n, 1), dtype=torch.uint8, pin_memory=True)
    return buf[:n].view(shape)


def _plan_nets(opt, seg_weights, mod_only, seg_only):
    """Which generators run (reference run_dask, models/__init__.py:296-325) and the keys of the result."""
    gens, segs = _names(opt)
    with_seg = bool(segs) and not mod_only
    w = _seg_weights(opt, seg_weights) if with_seg else None
    run_gens, run_segs = list(gens), (list(segs) if with_seg else None)
    if with_seg and seg_only:
        marker = f"G{opt.modalities_names.index('Marker')}" if "Marker" in opt.modalities_names else None
        run_segs = [k if w[i] != 0 else None for i, k in enumerate(segs)]
        run_gens = [k if (run_segs[i + 1] is not None or k == marker) else None for i, k in enumerate(gens)]
    return gens, segs, with_seg, w, run_gens, run_segs


def run_batch_device(tiles_dev, nets, opt, seg_weights=None, mod_only=False, micro_batch=8, n_streams=3, seg_only=False,
                     want_parts=True, variance=None):
    """uint8 tiles [T,ts,ts,3] ON THE DEVICE -> dict[name -> uint8 [T,ts,ts,3] on the device] with the reference's result
    keys (G1.., G{S}, and — with want_parts — the per-modality seg outputs G{S}k), skipping empty tiles exactly like
    run_wrapper (:399-443).  Nothing returns to the host here except the 3 integers per tile of the is_empty statistic.
    seg_only: seg generators whose weight is zero are not run, nor the modality generators that only feed them
    (models/__init__.py:318-325; the Marker generator always runs) — their keys are then absent from the result."""
    gens, segs, with_seg, w, run_gens, run_segs = _plan_nets(opt, seg_weights, mod_only, seg_only)
    T, ts = tiles_dev.shape[0], tiles_dev.shape[1]
    S = opt.mod_id_seg
    dev = tiles_dev.device
    keys = [k for k in run_gens if k] + ([f"G{S}"] + ([k for k in run_segs if k] if want_parts else []) if with_seg else [])
    out = {k: torch.zeros((T, ts, ts, 3), dtype=torch.uint8, device=dev) for k in keys}
    if T == 0:
        return out
    # is_empty(): gray-level variance per tile, computed on the device from the uint8 batch (exact integer sums)
    var = ops.tile_gray_variance(tiles_dev) if variance is None else variance
    live = [i for i in range(T) if var[i] >= EMPTY_TILE_VARIANCE]
    dead = sorted(set(range(T)) - set(live))
    if dead:
        di = torch.tensor(dead, device=dev)
        for j, k in enumerate(gens):
            if k in out:
                out[k][di] = torch.tensor(opt.background_colors[j], dtype=torch.uint8, device=dev)
    if live:
        all_live = len(live) == T
        li = None if all_live else torch.tensor(live, device=dev)
        batch = tiles_dev if all_live else tiles_dev.index_select(0, li)
        pipe = _pipeline(nets, run_gens, run_segs, w, micro_batch, n_streams)

        def put(key, val):
            if all_live:
                out[key].copy_(val)
            else:
                out[key].index_copy_(0, li, val)
        if with_seg:
            mods_u8, seg_u8, _, parts_u8 = pipe.infer_u8_device(batch, want_parts=want_parts)
            put(f"G{S}", seg_u8)
            if want_parts:
                for j, k in enumerate(pipe.part_index):
                    put(segs[k], parts_u8[j])
            for j, i in enumerate(pipe.mod_index):
                put(gens[i], mods_u8[j])
        else:
            x = ops.u8_to_f32(batch.contiguous())
            mb = micro_batch if micro_batch > 0 else x.shape[0]
            for i, k in enumerate(gens):
                o = torch.cat([nets[k](x[s_:s_ + mb]) for s_ in range(0, x.shape[0], mb)])
                put(k, ops.f32_to_u8(o))
    return out


def run_batch(tiles_u8, nets, opt, seg_weights=None, mod_only=False, micro_batch=8, n_streams=3, seg_only=False):
    """uint8 tiles [T,ts,ts,3] (numpy) -> dict[name -> uint8 numpy [T,ts,ts,3]]: host wrapper of run_batch_device (one
    pinned H2D copy in, one D2H copy per result key out)."""
    tiles_u8 = np.ascontiguousarray(tiles_u8)
    if tiles_u8.shape[0] == 0:
        dev_in = torch.zeros(tiles_u8.shape, dtype=torch.uint8, device="cuda")
    else:
        stage = _pinned("in", tiles_u8.shape)
        stage.copy_(torch.from_numpy(tiles_u8))
        dev_in = stage.to("cuda", non_blocking=True)
    res = run_batch_device(dev_in, nets, opt, seg_weights, mod_only, micro_batch, n_streams, seg_only)
    return {k: v.cpu().numpy() for k, v in res.items()}


def run_dask(img, model_path=None, nets=None, eager_mode=True, opt=None, seg_only=False, mod_only=False,
             seg_weights=None, use_dask=False, output_tensor=False):
    """Single-tile entry kept for API compatibility (name from the reference; nothing here uses Dask)."""
    assert model_path is not None or nets is not None, "Provide either the model path or the networks object."
    if nets is None:
        nets = init_nets(os.getenv("DEEPLIIF_MODEL_DIR", model_path), True, opt)
    tile = np.asarray(img.resize((opt.scale_size, opt.scale_size)).convert("RGB"))[None]
    res = run_batch(tile, nets, opt, seg_weights, mod_only, seg_only=seg_only)
    out = {k: Image.fromarray(v[0]) for k, v in res.items()}
    if seg_only:
        keep = [f"G{opt.mod_id_seg}", f"G{opt.modalities_no}"]
        out = {k: v for k, v in out.items() if k in keep}
    return out


def run_wrapper(tile, run_fn, model_path=None, nets=None, eager_mode=True, opt=None, seg_only=False, mod_only=False,
                seg_weights=None, use_dask=False, output_tensor=False):
    return run_fn(tile, model_path, nets, eager_mode, opt, seg_only, mod_only, seg_weights)


def infer_tiles(img, tile_size, overlap_size, nets, opt, seg_weights=None, mod_only=False, seg_only=False, micro_batch=8,
                n_streams=3, want_parts=True):
    """PIL image -> dict[net key -> stitched PIL image] (None on non-zero ranks of a torchrun launch).

    The tile -> infer -> stitch loop of the reference's inference() (models/__init__.py:484-500, InferenceTiler
    util/__init__.py:129-331).  One process per GPU: every rank tiles the image itself (cheap, deterministic), uploads and
    infers tiles rank, rank+W, ... (tile sharding, no data-path collective; SURVEY.md 8e, BASELINE config 3); the uint8
    results stay on the device, all output keys stacked, until ONE NCCL gather brings them to rank 0, which copies them to
    the host once and stitches.  want_parts=False skips the per-modality seg outputs (only `return_seg_intermediate` needs
    them): half the result bytes."""
    import torch.distributed as dist
    from .. import sharding
    grid = TileGrid(np.asarray(img.convert("RGB")), tile_size, overlap_size)
    tiles = grid.tiles()
    if tile_size != opt.scale_size:
        tiles = np.stack([np.asarray(Image.fromarray(t).resize((opt.scale_size, opt.scale_size))) for t in tiles])
    distributed = dist.is_available() and dist.is_initialized()      # a group of any size (also 1) takes the gather path
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    mine = np.ascontiguousarray(sharding.shard(tiles, rank, world)) if distributed else tiles
    if len(mine):
        stage = _pinned("in", mine.shape)
        stage.copy_(torch.from_numpy(mine))
        dev_in = stage.to("cuda", non_blocking=True)
    else:
        dev_in = torch.zeros((0,) + tuple(tiles.shape[1:]), dtype=torch.uint8, device="cuda")
    local = run_batch_device(dev_in, nets, opt, seg_weights, mod_only, micro_batch, n_streams, seg_only, want_parts)
    keys = sorted(local.keys())
    stacked = torch.stack([local[k] for k in keys], dim=1) if len(mine) else \
        torch.zeros((0, len(keys)) + tuple(tiles.shape[1:]), dtype=torch.uint8, device="cuda")
    if distributed:
        full = sharding.gather_to_rank0(stacked, len(tiles), to_numpy=False)       # [T, K, ts, ts, 3] on rank 0's device
        if rank != 0:
            return None
    else:
        full = stacked
    host = _pinned("out", tuple(full.shape))
    host.copy_(full, non_blocking=True)
    torch.cuda.synchronize()
    full_np = host.numpy()
    results = {}
    for j, k in enumerate(keys):
        v = full_np[:, j]
        if tile_size != opt.scale_size:
            v = np.stack([np.asarray(Image.fromarray(t).resize((tile_size, tile_size))) for t in v])
        results[k] = Image.fromarray(grid.stitch(v))
    return results


class _PendingImage:
    """Results of one image, still on their way: uint8 tiles in a pinned host buffer behind a CUDA event."""
    __slots__ = ("grid", "keys", "host", "event", "tile_size", "scale_size", "meta")


def _stitch_pending(pend):
    pend.event.synchronize()
    full_np = pend.host.numpy()
    results = {}
    for j, k in enumerate(pend.keys):
        v = full_np[:, j]
        if pend.tile_size != pend.scale_size:
            v = np.stack([np.asarray(Image.fromarray(t).resize((pend.tile_size, pend.tile_size))) for t in v])
        results[k] = Image.fromarray(pend.grid.stitch(v))
    return results


def infer_images(images, tile_size, overlap_size, nets, opt, seg_weights=None, mod_only=False, seg_only=False, micro_batch=8,
                 n_streams=3, want_parts=False, depth=3):
    """Pipelined `infer_tiles` over a sequence of PIL images (the per-image loop of `deepliif test`, cli.py:893-919, and
    the WSI sweep of BASELINE config 3): yields (index, dict[net key -> stitched PIL image]) in order (on non-zero ranks of
    a torchrun launch: (index, None)).

    Three things overlap instead of alternating: (i) the host tiles image k+1 (each rank only its own shard, straight
    into a pinned buffer) and uploads it on a side stream — where the is_empty statistic is also computed, so its small
    device-to-host read does not wait for the generators — while the GPU runs image k; (ii) the uint8 results of image k
    are gathered (one NCCL gather) and copied to a pinned buffer asynchronously; (iii) a stitcher thread on rank 0 turns
    finished buffers into images while the next ones are in flight.  `depth` buffers rotate."""
    import queue
    import threading
    import torch.distributed as dist
    from .. import sharding
    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    side = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    todo, done = queue.Queue(), queue.Queue()
    free = threading.Semaphore(depth)        # staging slots: taken by the main thread, returned when the image is stitched

    def stitcher():
        while True:
            item = todo.get()
            if item is None:
                return
            idx, pend = item
            try:
                done.put((idx, _stitch_pending(pend) if pend is not None else None))
            except Exception as e:           # surfaced by the consumer
                done.put((idx, e))
            free.release()

    th = threading.Thread(target=stitcher, daemon=True)
    th.start()
    n_sub = 0
    n_out = 0

    def drain(block):
        nonlocal n_out
        out = []
        while n_out < n_sub:
            try:
                idx, res = done.get(block=block and not out)
            except queue.Empty:
                break
            if isinstance(res, Exception):
                raise res
            out.append((idx, res)); n_out += 1
        return out

    slot = 0
    for idx, img in enumerate(images):
        free.acquire()                       # at most `depth` images between tiling and stitching: their buffers are distinct
        grid = TileGrid(np.asarray(img.convert("RGB")), tile_size, overlap_size)
        T = len(grid)
        mine_idx = sharding.shard_indices(T, rank, world)
        ts = grid.ts
        if tile_size != opt.scale_size:
            mine = grid.tiles(mine_idx)
            mine = np.stack([np.asarray(Image.fromarray(t).resize((opt.scale_size, opt.scale_size))) for t in mine]) \
                if len(mine) else np.zeros((0, opt.scale_size, opt.scale_size, 3), np.uint8)
            ts = opt.scale_size
        tag = ("in", slot % depth)
        if mine_idx:
            stage = _pinned(tag, (len(mine_idx), ts, ts, 3))
            if tile_size != opt.scale_size:
                stage.copy_(torch.from_numpy(np.ascontiguousarray(mine)))
            else:
                grid.tiles(mine_idx, out=stage.numpy())
            with torch.cuda.stream(side):
                dev_in = stage.to(dev, non_blocking=True)
                var = ops.tile_gray_variance(dev_in)              # syncs the SIDE stream only
            main.wait_stream(side)
            dev_in.record_stream(main)
        else:
            dev_in = torch.zeros((0, ts, ts, 3), dtype=torch.uint8, device=dev)
            var = np.zeros((0,))
        local = run_batch_device(dev_in, nets, opt, seg_weights, mod_only, micro_batch, n_streams, seg_only, want_parts,
                                 variance=var)
        keys = sorted(local.keys())
        stacked = torch.stack([local[k] for k in keys], dim=1) if mine_idx else \
            torch.zeros((0, len(keys), ts, ts, 3), dtype=torch.uint8, device=dev)
        full = sharding.gather_to_rank0(stacked, T, to_numpy=False) if distributed else stacked
        pend = None
        if rank == 0:
            pend = _PendingImage()
            pend.grid, pend.keys, pend.tile_size, pend.scale_size = grid, keys, tile_size, opt.scale_size
            pend.host = _pinned(("out", slot % depth), tuple(full.shape))
            pend.host.copy_(full, non_blocking=True)
            pend.event = torch.cuda.Event()
            pend.event.record(main)
        todo.put((idx, pend))
        n_sub += 1
        slot += 1
        for item in drain(block=False):
            yield item
    todo.put(None)
    for item in drain(block=True):
        yield item
    while n_out < n_sub:
        for item in drain(block=True):
            yield item
    th.join()


def inference(img, tile_size, overlap_size, model_path, use_torchserve=False, eager_mode=True, color_dapi=False,
              color_marker=False, opt=None, return_seg_intermediate=False, seg_only=False, mod_only=False,
              seg_weights=None, opt_args={}):
    """PIL image -> dict[name -> PIL image] with the reference's output names (mod{i}-{Name}, Seg, {mod}_s)."""
    if use_torchserve:
        raise NotImplementedError("TorchServe is outside the B200 hot-path scope")
    if opt is None:
        opt = get_opt(model_path)
    for k, v in opt_args.items():
        setattr(opt, k, v)
    if getattr(opt, "seg_gen", True) is False and (seg_only or return_seg_intermediate):      # models/__init__.py:478-482
        seg_only = return_seg_intermediate = False
        print("option seg_gen is False, disabled seg_only and return_seg_intermediate")
    # seg_weights=None means equal weights 1/(modalities_no+1), as in the reference's run_dask (:299-306): only the
    # `deepliif test` command passes opt.seg_weights down (cli.py:878, 906), a direct API call does not read them
    if getattr(opt, "input_no", 1) > 1:
        raise NotImplementedError("inference(): models with several input images side by side (input_no > 1, SDG-style) are "
                                  "outside the B200 hot-path scope")
    nets = init_nets(os.getenv("DEEPLIIF_MODEL_DIR", model_path), True, opt)
    results = infer_tiles(img, tile_size, overlap_size, nets, opt, seg_weights=seg_weights, mod_only=mod_only, seg_only=seg_only,
                          want_parts=bool(return_seg_intermediate) and not seg_only)
    if results is None:
        return {}
    # ---- the reference's naming (models/__init__.py:502-565) -----------------------------------------------
    n, S = opt.modalities_no, opt.mod_id_seg
    names = opt.modalities_names
    mod_names = [f"mod{i + 1}" for i in range(n)]
    if mod_names != names[opt.input_no:]:
        mod_names = [f"mod{i + 1}-{nm}" for i, nm in enumerate(names[opt.input_no:])]
    name2id = {nm: f"G{i + 1}" for i, nm in enumerate(mod_names)}
    if not mod_only and opt.seg_gen:
        name2id["Seg"] = f"G{S}"
    if seg_only:
        images = {"Seg": results[name2id["Seg"]]}
        marker = [k for k in name2id if k.endswith("Marker")]
        if marker:
            images[marker[0]] = results[name2id[marker[0]]]
        return images
    images = {nm: results[mid] for nm, mid in name2id.items()}
    if opt.seg_gen and return_seg_intermediate and not mod_only:
        seg_names = [f"mod{i}" for i in range(n + 1)]
        if seg_names != names:
            seg_names = [f"mod{i}-{nm}" for i, nm in enumerate(names)]
        base = 0 if f"G{S}0" in results else 1
        images.update({f"{nm}_s": results[f"G{S}{i + base}"] for i, nm in enumerate(seg_names)})
    return images


def find_marker_key(dictionary):
    """models/__init__.py:950-954."""
    for key in dictionary:
        if key.endswith("Marker"):
            return key
    return None


def postprocess(orig, images, tile_size, model, seg_thresh=120, size_thresh="default", marker_thresh=None,
                size_thresh_upper=None):
    """Reference models/__init__.py:582-610: cell-level scoring + SegOverlaid / SegRefined from the stitched Seg (and
    Marker) image; the pixel and graph work runs on the GPU (deepliif_b200.postprocessing, csrc/cells.cu)."""
    from ..postprocessing import compute_final_results
    if model in ("DeepLIIF", "DeepLIIFKD"):
        resolution = "40x" if tile_size > 384 else ("20x" if tile_size > 192 else "10x")
        marker_key = find_marker_key(images)
        overlay, refined, scoring = compute_final_results(
            orig, images["Seg"], images.get(marker_key) if marker_key else None, resolution, size_thresh, marker_thresh,
            size_thresh_upper, seg_thresh)
        return {"SegOverlaid": Image.fromarray(overlay), "SegRefined": Image.fromarray(refined)}, scoring
    raise Exception(f"postprocess() not implemented for model {model}")


def infer_modalities(img, tile_size, model_dir, eager_mode=True, color_dapi=False, color_marker=False, opt=None,
                     return_seg_intermediate=False, seg_only=False, mod_only=False, seg_weights=None):
    if opt is None:
        opt = get_opt(model_dir)
    if not tile_size:
        tile_size = opt.scale_size
    images = inference(img, tile_size=tile_size, overlap_size=tile_size // 16, model_path=model_dir,
                       eager_mode=True, color_dapi=color_dapi, color_marker=color_marker, opt=opt,
                       return_seg_intermediate=return_seg_intermediate, seg_only=seg_only, mod_only=mod_only,
                       seg_weights=seg_weights)
    if not images:                                  # non-zero ranks of a tile-sharded run: rank 0 holds the results
        return images, None
    if getattr(opt, "seg_gen", True) and not mod_only:          # models/__init__.py:648-657
        post_images, scoring = postprocess(img, images, tile_size, opt.model)
        images = {**images, **post_images}
        if seg_only:
            images = {k: v for k, v in images.items() if "Seg" in k}
        return images, scoring
    return images, None
