"""Sharded rendering of ONE frame over the GPUs of a box.

Host layer of csrc/shard.cu: one process per GPU (torchrun), torch.distributed (NCCL over
NVLink 5 / NVSwitch) for the plumbing -- the G x G count matrix, the barrier after the peer-memory
exchange and the gather of the finished bands -- while the splats themselves move by direct stores
into the owners' buffers from inside the exchange kernel.  The reference is single-GPU; nothing here
has an upstream counterpart beyond GaussianRenderer::prepare/render (renderer.rs:191-260).
"""
import ctypes as C

import numpy as np


def tile_row_bands(height, world):
    """rank d owns tile rows [b[d], b[d+1]) of the ceil(height/16) rows -- mirrors ws_renderer_shard_configure."""
    ty = (height + 15) // 16
    return [(ty * d) // world for d in range(world + 1)]


def balanced_bands(bands, loads, min_rows=1):
    """Cost-balanced tile-row bands from the per-band load of the last frame(s) ("a cost-balanced
    map using the gathered counts").  The load is taken as uniform over the rows of its band, which gives a piecewise-
    linear cumulative cost over the tile rows; the new boundaries cut it into equal parts, rounded to whole rows with at
    least `min_rows` rows per rank.  Pure host arithmetic on values every rank holds identically, so every rank derives
    the same bands.  Applying it a few times converges (each round refines the density estimate)."""
    world = len(bands) - 1
    ty = int(bands[-1])
    loads = [max(float(x), 0.0) for x in loads]
    total = sum(loads)
    if total <= 0.0 or ty < world * min_rows:
        return [int(b) for b in bands]
    cum = [0.0] * (ty + 1)                                  # cumulative cost at every row boundary
    for d in range(world):
        rows = bands[d + 1] - bands[d]
        for y in range(bands[d], bands[d + 1]):
            cum[y + 1] = cum[y] + loads[d] / rows
    out = [0]
    for d in range(1, world):
        target = total * d / world
        y = out[-1] + min_rows
        hi = ty - (world - d) * min_rows                    # leave room for the ranks behind
        while y < hi and abs(cum[y + 1] - target) <= abs(cum[y] - target):
            y += 1
        out.append(min(y, hi))
    out.append(ty)
    return out


def shard_cloud(cloud, rank, world):
    """Gaussians [rank*N/world, (rank+1)*N/world) of a cloud dict, keeping the GLOBAL bbox / centre
    (clip box, scene centre and extent come from them: renderer.rs:622-651)."""
    n = int(cloud["num_points"])
    lo, hi = (n * rank) // world, (n * (rank + 1)) // world
    out = dict(cloud)
    out["gaussians"] = cloud["gaussians"][lo:hi]
    if cloud["compressed"]:
        out["sh_coefs"] = cloud["sh_coefs"]              # codebooks are replicated
    else:
        out["sh_coefs"] = cloud["sh_coefs"][lo:hi]
    out["num_points"] = hi - lo
    out["first_index"] = lo
    return out


def exchange_offsets(matrix, rank):
    """Where rank `rank`'s records start in each destination (host mirror of route_scatter_kernel)."""
    m = np.asarray(matrix)
    return m[:rank].sum(axis=0)


class ShardedRenderer:
    """GaussianRenderer over `world` GPUs.  Every rank calls frame() with the same SplattingArgs;
    the full frame is returned on every rank (device tensor) after the band all-gather."""

    def __init__(self, ws, ctx, color_format, sh_deg, compressed, pc, total_points, viewport, group=None,
                 pair_capacity=None):
        import torch
        import torch.distributed as dist
        self.ws, self.torch, self.dist = ws, torch, dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.pc = pc
        self.W, self.H = int(viewport[0]), int(viewport[1])
        self.format = int(color_format)
        self.r = ws.GaussianRenderer.new(ctx, color_format, sh_deg, compressed)
        if pair_capacity:
            self.r.set_pair_capacity(pair_capacity)
        L = ws.lib()
        ws._check(L.ws_renderer_shard_configure(self.r._h, self.rank, self.world, int(total_points), pc.num_points(), self.W, self.H))
        self.r._viewport = (self.W, self.H)
        handles = torch.zeros(384, dtype=torch.uint8)
        ws._check(L.ws_renderer_shard_export(self.r._h, C.c_void_p(handles.data_ptr())))
        if self.world > 1:
            dev = torch.device("cuda", torch.cuda.current_device())
            allh = [torch.zeros(384, dtype=torch.uint8, device=dev) for _ in range(self.world)]
            dist.all_gather(allh, handles.to(dev), group=group)
            allh = torch.cat([h.cpu() for h in allh]).contiguous()
            ws._check(L.ws_renderer_shard_import(self.r._h, C.c_void_p(allh.data_ptr())))
        dev = torch.device("cuda", torch.cuda.current_device())
        self.row = torch.zeros(self.world, dtype=torch.int32, device=dev)
        self.matrix = torch.zeros(self.world * self.world, dtype=torch.int32, device=dev)
        self.flag = torch.zeros(1, dtype=torch.int32, device=dev)
        first, rows = C.c_uint32(), C.c_uint32()
        ws._check(L.ws_renderer_shard_band(self.r._h, C.byref(first), C.byref(rows)))
        self.first_row, self.num_rows = first.value, rows.value
        self.bands = tile_row_bands(self.H, self.world)
        self.max_rows = max(min(self.bands[d + 1] * 16, self.H) - min(self.bands[d] * 16, self.H) for d in range(self.world))
        dt = {0: torch.uint8, 1: torch.float16, 2: torch.float32}[self.format]
        self.band = torch.zeros((self.max_rows, self.W, 4), dtype=dt, device=dev)
        self.gathered = torch.zeros((self.world, self.max_rows, self.W, 4), dtype=dt, device=dev)
        self.frame_out = torch.zeros((self.H, self.W, 4), dtype=dt, device=dev)
        self._frames = 1                                         # mirrors the library's epoch (first frame = 1)
        self._copied = [None, None]

    def set_bands(self, bands):
        """Install custom tile-row bands (same list on every rank, no frame in flight)."""
        torch, ws = self.torch, self.ws
        arr = (C.c_uint32 * (self.world + 1))(*[int(b) for b in bands])
        ws._check(ws.lib().ws_renderer_shard_set_bands(self.r._h, arr, self.world + 1))
        self.bands = [int(b) for b in bands]
        first, rows = C.c_uint32(), C.c_uint32()
        ws._check(ws.lib().ws_renderer_shard_band(self.r._h, C.byref(first), C.byref(rows)))
        self.first_row, self.num_rows = first.value, rows.value
        max_rows = max(min(self.bands[d + 1] * 16, self.H) - min(self.bands[d] * 16, self.H) for d in range(self.world))
        if max_rows != self.max_rows:                            # buffers of the all-gather variant
            self.max_rows = max_rows
            self.band = torch.zeros((max_rows, self.W, 4), dtype=self.band.dtype, device=self.band.device)
            self.gathered = torch.zeros((self.world, max_rows, self.W, 4), dtype=self.band.dtype, device=self.band.device)

    def band_loads(self):
        """Pairs each rank produced in its band in the last frame (all ranks get the same list)."""
        torch, dist = self.torch, self.dist
        torch.cuda.synchronize()
        mine = torch.tensor([float(self.r.stats(allow_overflow=True)["num_pairs"])], dtype=torch.float64, device="cuda")
        if self.world == 1:
            return [float(mine.item())]
        allv = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(allv, mine, group=self.group)
        return [float(v.item()) for v in allv]

    def rebalance(self):
        """Re-cut the bands from the last frame's per-band pair counts; returns the new bands."""
        new = balanced_bands(self.bands, self.band_loads())
        if new != self.bands:
            self.set_bands(new)
        return new

    def frame(self, args, clear=(0.0, 0.0, 0.0, 0.0), gather=True, marks=None):
        """Enqueue one frame on torch's current stream; returns the assembled frame (device tensor).
        `marks` (optional list) receives (label, torch.cuda.Event) pairs between the phases."""
        torch, dist, ws = self.torch, self.dist, self.ws
        L = ws.lib()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def mark(label):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((label, e))

        a = args._c()
        mark("start")
        ws._check(L.ws_renderer_shard_begin(self.r._h, self.pc._h, C.byref(a), C.c_void_p(self.row.data_ptr()), stream))
        mark("stage1+route_count")
        if self.world > 1:
            dist.all_gather_into_tensor(self.matrix, self.row, group=self.group)            # G x G counts
        else:
            self.matrix.copy_(self.row)
        mark("allgather_counts")
        ws._check(L.ws_renderer_shard_exchange(self.r._h, C.c_void_p(self.matrix.data_ptr()), stream))
        mark("exchange_kernel")
        if self.world > 1:
            dist.all_reduce(self.flag, group=self.group)                                       # every rank's stores have landed
        mark("barrier")
        ws._check(L.ws_renderer_shard_finish(self.r._h, C.c_void_p(self.matrix.data_ptr()), stream))
        mark("sort+binning")
        clr = (C.c_double * 4)(*[float(c) for c in clear])
        pitch = self.W * ws._BPP[self.format]
        ws._check(L.ws_renderer_render_band(self.r._h, self.pc._h, C.c_void_p(self.band.data_ptr()), pitch, C.byref(clr), stream))
        mark("composite_band")
        if not gather:
            return self.band[: self.num_rows]
        if self.world > 1:
            dist.all_gather_into_tensor(self.gathered, self.band, group=self.group)
            mark("allgather_bands")
            for d in range(self.world):
                y0, y1 = min(self.bands[d] * 16, self.H), min(self.bands[d + 1] * 16, self.H)
                if y1 > y0:
                    self.frame_out[y0:y1].copy_(self.gathered[d, : y1 - y0])
        else:
            self.frame_out.copy_(self.band[: self.H])
        mark("assemble")
        return self.frame_out

    def frame_to_root(self, args, clear=(0.0, 0.0, 0.0, 0.0), root=0, host=None, marks=None):
        """One frame whose bands are stored directly into rank `root`'s assembled frame (peer memory);
        no image collective, only a barrier.  If `host` (pinned CPU tensor, root only) is given the
        assembled frame is downloaded into it asynchronously.  Returns nothing: read the frame on the
        root with download()/ws_renderer_shard_frame after synchronising."""
        torch, dist, ws = self.torch, self.dist, self.ws
        L = ws.lib()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def mark(label):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((label, e))

        a = args._c()
        mark("start")
        ws._check(L.ws_renderer_shard_begin(self.r._h, self.pc._h, C.byref(a), C.c_void_p(self.row.data_ptr()), stream))
        mark("stage1+route_count")
        if self.world > 1:
            dist.all_gather_into_tensor(self.matrix, self.row, group=self.group)
        else:
            self.matrix.copy_(self.row)
        mark("allgather_counts")
        ws._check(L.ws_renderer_shard_exchange(self.r._h, C.c_void_p(self.matrix.data_ptr()), stream))
        mark("exchange_kernel")
        if self.world > 1:
            dist.all_reduce(self.flag, group=self.group)
        mark("barrier")
        ws._check(L.ws_renderer_shard_finish(self.r._h, C.c_void_p(self.matrix.data_ptr()), stream))
        mark("sort+binning")
        clr = (C.c_double * 4)(*[float(c) for c in clear])
        ws._check(L.ws_renderer_render_band_to_root(self.r._h, self.pc._h, root, C.byref(clr), stream))
        mark("composite_band->root")
        if self.world > 1:
            dist.all_reduce(self.flag, group=self.group)                 # all bands have landed in the root's frame
        mark("barrier2")
        if host is not None and self.rank == root:
            ws._check(L.ws_renderer_shard_download(self.r._h, C.c_void_p(host.data_ptr()), stream))

    def frame_peer(self, args, clear=(0.0, 0.0, 0.0, 0.0), root=0, host=None, copy_stream=None):
        """One frame with NO host-side collective at all (ws_renderer_shard_frame_to_root): a single C
        call per rank; rows, barrier and band arrival are flags in peer-mapped mailboxes.  With `host`
        (root only) the assembled frame is downloaded; with `copy_stream` the download runs there and
        overlaps the next frame (the root keeps two frame buffers, alternating per frame)."""
        ws, torch = self.ws, self.torch
        L = ws.lib()
        cur = torch.cuda.current_stream()
        stream = C.c_void_p(cur.cuda_stream)
        par = self._frames & 1                                   # parity of the frame buffer this frame fills
        self._frames += 1
        is_root = self.rank == root
        if is_root and copy_stream is not None and self._copied[par] is not None:
            cur.wait_event(self._copied[par])                    # its previous download must be over before peers overwrite it
        a = args._c()
        clr = (C.c_double * 4)(*[float(c) for c in clear])
        ws._check(L.ws_renderer_shard_frame_to_root(self.r._h, self.pc._h, C.byref(a), root, C.byref(clr), stream))
        if host is not None and is_root:
            if copy_stream is None:
                ws._check(L.ws_renderer_shard_download(self.r._h, C.c_void_p(host.data_ptr()), stream))
            else:
                done = torch.cuda.Event()
                done.record(cur)
                copy_stream.wait_event(done)
                ws._check(L.ws_renderer_shard_download(self.r._h, C.c_void_p(host.data_ptr()), C.c_void_p(copy_stream.cuda_stream)))
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                self._copied[par] = ev

    def download(self, host):
        """root only: asynchronous copy of the assembled frame into `host` (pinned CPU tensor / numpy array)."""
        ptr = host.ctypes.data if isinstance(host, np.ndarray) else host.data_ptr()
        stream = C.c_void_p(self.torch.cuda.current_stream().cuda_stream)
        self.ws._check(self.ws.lib().ws_renderer_shard_download(self.r._h, C.c_void_p(ptr), stream))

    def stats(self, allow_overflow=False):
        return self.r.stats(allow_overflow)


class ShardedPipeline:
    """`depth` sharded frames in flight per GPU: one ShardedRenderer (own scratch, mailboxes, peer mappings, frame
    buffers) and one CUDA stream per frame slot, all sharing the resident point-cloud shard.  Frame i runs in slot
    i % depth; every rank must submit the same frames in the same order.  At small per-GPU shares the kernels of one
    frame are latency-bound and leave most SMs idle; a second frame fills them (and the waits on peer flags).  With
    depth > 1 the flag waits run in one-warp gate kernels (ws_renderer_shard_set_gated) so that two frames can never
    starve each other across GPUs."""

    def __init__(self, ws, ctx, color_format, sh_deg, compressed, pc, total_points, viewport, depth=2, group=None,
                 pair_capacity=None):
        import torch
        self.torch, self.ws, self.depth = torch, ws, int(depth)
        self.slots = [ShardedRenderer(ws, ctx, color_format, sh_deg, compressed, pc, total_points, viewport, group, pair_capacity)
                      for _ in range(self.depth)]
        self.streams = [torch.cuda.Stream() for _ in range(self.depth)]
        self.rank, self.world = self.slots[0].rank, self.slots[0].world
        for s in self.slots:
            s.r.set_timing(False)
            ws._check(ws.lib().ws_renderer_shard_set_gated(s.r._h, 1 if self.depth > 1 else 0))
        self._i = 0

    def frame_peer(self, args, clear=(0.0, 0.0, 0.0, 0.0), root=0, host=None, copy_stream=None):
        k = self._i % self.depth
        self._i += 1
        with self.torch.cuda.stream(self.streams[k]):
            self.slots[k].frame_peer(args, clear, root, host, copy_stream)
        return k

    def synchronize(self):
        for st in self.streams:
            st.synchronize()

    @property
    def bands(self):
        return self.slots[0].bands

    def rebalance(self):
        """Balance from slot 0's last frame and install the same bands in every slot."""
        self.synchronize()
        new = balanced_bands(self.slots[0].bands, self.slots[0].band_loads())
        for s in self.slots:
            if new != s.bands:
                s.set_bands(new)
        return new
