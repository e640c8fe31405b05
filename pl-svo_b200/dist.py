"""Batch sharding across the GPUs of one box.

Frame pairs are independent units: rank r of G owns a contiguous slice of the batch, runs the
alignment kernel on it with no data-path collective, and the per-pair results (7-double poses,
< 0.5 KB per pair with H) are gathered with one all_gather.  torch.distributed is the plumbing
(NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [begin, end) slice of rank `rank`; the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_rows(local: np.ndarray, n_items: int, device: torch.device | str = "cpu") -> np.ndarray:
    """all_gather of per-item result rows (e.g. [n_local, 7] poses) into the global [n_items, ...] array,
    in batch order, on every rank.  Uneven shards are padded to the largest shard for the collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    unsigned = {np.dtype(np.uint16): np.int16, np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}
    if local.dtype in unsigned:  # the collectives carry signed integers; same bits
        back = local.dtype
        return gather_rows(np.ascontiguousarray(local).view(unsigned[local.dtype]), n_items, device).view(back)
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    cap = max(e - b for b, e in sizes)
    row = local.shape[1:]
    buf = torch.zeros((cap,) + row, dtype=torch.from_numpy(local).dtype, device=device)
    buf[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(device)
    out = torch.empty((world * cap,) + row, dtype=buf.dtype, device=device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape((world, cap) + row)
    return np.concatenate([out[r, : e - b] for r, (b, e) in enumerate(sizes)], axis=0)


# ------------------------------------------------------------------------------------------------
# One host batch on G GPUs: scatter -> align -> gather
# ------------------------------------------------------------------------------------------------
_ALIGN_ARRAYS = ("T_ref_w", "T_cur_w", "T_cur_w_gt", "pt_px", "pt_f", "pt_pos", "seg_spx", "seg_epx", "seg_sf", "seg_ef", "seg_spos",
                 "seg_epos", "seg_length", "pt_valid", "seg_valid", "pt_count", "seg_count", "pt_depth", "seg_sdepth", "seg_edepth")
_RESULT_FIELDS = ("T_cur_w", "n_tracked", "H", "seg_killed", "iters", "status", "patch_iters", "patch_levels")


_pinned = {}
_thread_pool = None


def _pool():
    """A process-wide pool of a few host threads for the packing copies (created once, not per call)."""
    global _thread_pool
    if _thread_pool is None:
        from concurrent.futures import ThreadPoolExecutor

        _thread_pool = ThreadPoolExecutor(max_workers=8)
    return _thread_pool


def _staging(key, nbytes, pin):
    """A reusable (pinned when CUDA is in play) uint8 staging tensor of at least nbytes."""
    t = _pinned.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(max(nbytes, 256), dtype=torch.uint8, pin_memory=pin)
        _pinned[key] = t
    return t


def _manifest(arrays):
    manifest, off = [], 0
    for name, a in arrays:
        manifest.append((name, str(a.dtype), a.shape, off, a.nbytes))
        off = (off + a.nbytes + 255) // 256 * 256
    return manifest, max(off, 256)


def _pack_jobs(buf, arrays, manifest):
    """One copy job per array: to its 256-byte aligned place inside the uint8 block `buf` (a numpy view)."""
    def job(o, nb, a):
        def run():
            buf[o:o + nb] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        return run
    return [job(o, nb, a) for (name, dt, shape, o, nb), (_, a) in zip(manifest, arrays)]


def _unpack(buf, manifest):
    return {name: buf[o:o + nb].view(np.dtype(dt)).reshape(shape) for name, dt, shape, o, nb in manifest}


def _shard_arrays(data, b, e):
    out = []
    for name in _ALIGN_ARRAYS:
        a = getattr(data, name, None)
        if a is not None:
            out.append((name, a[b:e]))
    for l, v in data.ref_pyr.items():
        out.append((f"ref_pyr/{l}", v[b:e]))
    for l, v in data.cur_pyr.items():
        out.append((f"cur_pyr/{l}", v[b:e]))
    return out


def align_sharded(data, max_level: int = 4, min_level: int = 2, n_iter: int = 30, src: int = 0, device=None, run_fn=None, ctx=None):
    """plsvo::SparseImgAlign::run over ONE host batch that lives on rank `src`, on all GPUs of the process group:
    rank `src` cuts the batch into contiguous shards (shard_range), packs each shard's inputs into one block and
    scatters the blocks (NCCL over NVLink on the GPUs, gloo in the CPU tests); every rank aligns its shard; the per-pair
    results come back with one all_gather per output, in batch order, on every rank.  Pairs are independent: there is no
    collective on the data path between the scatter and the gather.

    data: an AlignData on rank `src` (ignored elsewhere, may be None).  run_fn(shard AlignData) -> AlignOut defaults to the
    CUDA path (SparseImgAlign.run through the C ABI, on `ctx` or on a context of this rank's device); the CPU test
    injects its own.  Returns a dict of result arrays."""
    from . import synth

    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = (run_fn or _default_run(max_level, min_level, n_iter, ctx, device))(data)
        return {f: getattr(out, f) for f in _RESULT_FIELDS}
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = torch.device(device) if device is not None else torch.device("cpu")
    # 1. what every rank needs to know before the scatter: shard sizes, manifests, camera
    meta = [None]
    bufs = None
    pin = dev.type == "cuda"
    if rank == src:
        n = data.batch
        shards = [_shard_arrays(data, *shard_range(n, r, world)) for r in range(world)]
        mans = [_manifest(sh) for sh in shards]
        cap = max(sz for _, sz in mans)
        meta = [{"n": n, "cap": cap, "manifests": [mf for mf, _ in mans],
                 "cam": (data.cam.width, data.cam.height, data.cam.fx, data.cam.fy, data.cam.cx, data.cam.cy),
                 "levels": (data.max_level, data.min_level)}]
        stage = _staging("src", cap * world, pin)  # one pinned block, one H2D copy for all shards
        host = stage.numpy()
        jobs = []
        for r, (sh, (mf, _)) in enumerate(zip(shards, mans)):
            jobs += _pack_jobs(host[r * cap:(r + 1) * cap], sh, mf)
        # the packing is plain memory copies (NumPy releases the GIL): spread them over a few host threads
        list(_pool().map(lambda j: j(), jobs))
        dev_all = stage[: cap * world].to(dev, non_blocking=True)
        bufs = list(dev_all.split(cap))
    dist.broadcast_object_list(meta, src=src)
    m = meta[0]
    # 2. scatter of the packed shards
    mine = torch.empty(m["cap"], dtype=torch.uint8, device=dev)
    dist.scatter(mine, bufs if rank == src else None, src=src)
    back = _staging("dst", m["cap"], pin)  # the C ABI reads host buffers: pinned, so its own H2D runs at full PCIe speed
    back[: m["cap"]].copy_(mine)
    if pin:
        torch.cuda.synchronize(dev)
    arrays = _unpack(back.numpy(), m["manifests"][rank])
    b, e = shard_range(m["n"], rank, world)
    shard = synth.AlignData(
        cam=synth.Camera(*m["cam"]), max_level=m["levels"][0], min_level=m["levels"][1],
        ref_pyr={int(k.split("/")[1]): np.ascontiguousarray(v) for k, v in arrays.items() if k.startswith("ref_pyr/")},
        cur_pyr={int(k.split("/")[1]): np.ascontiguousarray(v) for k, v in arrays.items() if k.startswith("cur_pyr/")},
        **{k: np.ascontiguousarray(arrays[k]) if k in arrays else None for k in _ALIGN_ARRAYS[:13]})
    for k in _ALIGN_ARRAYS[13:]:
        if k in arrays:
            setattr(shard, k, np.ascontiguousarray(arrays[k]))
    # 3. every rank aligns its shard (empty shards run nothing)
    res = {}
    if e > b:
        out = (run_fn or _default_run(max_level, min_level, n_iter, ctx, dev))(shard)
        res = {f: np.asarray(getattr(out, f)) for f in _RESULT_FIELDS}
    # 4. gather, in batch order, on every rank (row shapes are fixed by the ABI: include/plsvo_b200.h plsvo_align_result)
    from . import abi

    n_segs = shard.n_segs
    spec = {"T_cur_w": (np.float64, (7,)), "n_tracked": (np.int64, ()), "H": (np.float64, (36,)), "seg_killed": (np.uint8, (n_segs,)),
            "iters": (np.int32, (abi.MAX_LEVELS,)), "status": (np.int32, ()), "patch_iters": (np.uint32, ()), "patch_levels": (np.uint32, ())}
    # one record per pair holding every output field, so that the results travel in ONE all_gather
    widths = [int(np.dtype(spec[f][0]).itemsize * int(np.prod(spec[f][1], dtype=np.int64))) for f in _RESULT_FIELDS]
    rec = np.zeros((e - b, sum(widths)), np.uint8)
    off = 0
    for f, w in zip(_RESULT_FIELDS, widths):
        dt, row = spec[f]
        if e > b:
            rec[:, off:off + w] = np.ascontiguousarray(res[f].astype(dt, copy=False)).reshape(e - b, -1).view(np.uint8)
        off += w
    allrec = gather_rows(rec, m["n"], device=dev)
    full, off = {}, 0
    for f, w in zip(_RESULT_FIELDS, widths):
        dt, row = spec[f]
        full[f] = np.ascontiguousarray(allrec[:, off:off + w]).view(dt).reshape((m["n"],) + row)
        off += w
    return full


_ctx_cache = {}


def _default_run(max_level, min_level, n_iter, ctx=None, device=None):
    from .api import Context, SparseImgAlign

    if ctx is None:
        index = torch.device(device).index if device is not None and torch.device(device).type == "cuda" else 0
        index = 0 if index is None else index
        if index not in _ctx_cache:
            _ctx_cache[index] = Context(index)
        ctx = _ctx_cache[index]

    def run(shard):
        return SparseImgAlign(max_level, min_level, n_iter, ctx=ctx).run(shard)

    return run
