"""FrontEnd — the data-parallel front end of the rotation-invariant PVCNN feature extractor as ONE engine.

One "step" takes a batch of clouds (xyz | normal, [B,6,N]) and their per-point input features ([B,C,N]) and
produces everything the dense layers of a PVConv block consume (SURVEY.md §3 call stack A, §8 rows a1-a10):

    branch A (ALU bound)      k-NN self query  ->  fused neighbour gather + PPF          -> ppf   [B,4,k,N]
    branch B (HBM bound)      coordinate prologue (torch mean + one kernel) -> voxelize        -> grid  [B,C,r,r,r], ind, cnt
                              -> trilinear devoxelize of the grid                        -> devox [B,C,N]
                              (the voxelizer emits the DGCNN voxel-neighbour edge features -> edge [B,2C,N]
                               in the same pass: it already holds every point's cell mean)

The two branches are independent, so they run on two streams inside one captured CUDA graph: the brute-force
k-NN (register/ALU bound) overlaps the dense grid write (HBM bound).  All buffers are allocated once; the kernels
are enqueued straight through the C ABI (no per-op allocation, no dispatcher overhead) and replayed as a graph.

`forward()`          runs one step on the device-resident input buffers (what bench.py's `value` times).
`__call__(pts, feats)`  is the host-facing call: pinned host inputs -> H2D -> step -> D2H of the per-point
                        outputs into pinned host buffers (what bench.py's `e2e` times).
"""
import torch

from . import _lib

_L = _lib.lib
_check = _lib.check


class FrontEnd:
    # knn+ppf (fused; else split, knn, ppf_gather), vox_prologue, vox_prepare + per grid chunk: vox_means (+edge
    # features), vox_fill, devox
    @property
    def kernels_per_step(self):
        return 3 + (3 if self._fused_front else 2 + 3 * self.grid_chunks)
    NORM_MODE = 1            # association of the 3-term radius sum that matches torch's norm kernel (see prologue.cu)

    L2_CHUNK_BYTES = 40 << 20   # dense-grid bytes per pipeline chunk: small enough to still be in L2 when consumed

    def __init__(self, B, N, C, k=20, r=32, voxel_shape='spherical', normalize=False, eps=0.0,
                 device='cuda', use_graph=True, overlap=True, grid_chunks=None, devox_side_stream=True,
                 knn_after_front=True, join_before_devox=None, fuse_mean=True, edge_echo=True):
        if voxel_shape not in ('spherical', 'cube'):
            raise ValueError('voxel_shape must be "spherical" or "cube"')
        self.B, self.N, self.C, self.k, self.r = int(B), int(N), int(C), int(k), int(r)
        self.voxel_shape, self.normalize, self.eps = voxel_shape, normalize, eps
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('FrontEnd runs on a CUDA device only (no CPU path exists)')
        self.use_graph, self.overlap = use_graph, overlap
        # edge_echo=False: `edge` is [B,C,N] = feat - mean(cell) only.  The other half of the DGCNN edge tensor (pvconv.py:90:
        # cat(rel, feat)) is the input itself; a host-side caller that just sent the features over PCIe does not want them
        # back.  Only the fused prefix kernel has this mode.
        self.edge_echo = bool(edge_echo)
        B, N, C, k, r = self.B, self.N, self.C, self.k, self.r
        s = r ** 3
        # The grid goes through L2 in chunks of clouds: voxelize(chunk) -> devoxelize(chunk) while the chunk's dense
        # grid (C * r^3 * 4 bytes per cloud) is still cache-resident, the next chunk's fill overlapping this chunk's
        # devoxelize on a second stream.  Chunking needs the two-phase voxelizer (tiled path: N <= 4096, r^3 % 4 == 0).
        per_cloud = max(1, (C + 1) * s * 4)
        if grid_chunks is None:
            # measured on B200 (tools/tune_step.py): the fixed cost of the extra launches outweighs the L2 hits at
            # 32 x 1024-point clouds (1 chunk 268 us, 2 chunks 282 us, 4 chunks 295 us), so one chunk is the default
            grid_chunks = 1
        if N > 4096 or s % 4 != 0:
            grid_chunks = 1
        self.grid_chunks = int(max(1, min(grid_chunks, max(B, 1))))
        self.devox_side_stream = bool(devox_side_stream)
        if join_before_devox is None:
            # The cube devoxelizer streams the grid through shared memory (devox.cu, streaming form) and belongs to the same
            # L1/shared-memory split as the k-NN and the grid writer, so it starts as soon as the grid is written
            # (measured step 192 -> 179 us).  The spherical one (gather form) only touches the ~40 cells per plane its index
            # quirks can reach, so its L1 appetite does not matter either (165 -> 158 us without the join).  Only the cube
            # gather form (r = 64, N > 2048, odd r) lives off a large L1 and waits for the k-NN to leave the SMs.
            join_before_devox = voxel_shape == 'cube' and (N > 2048 or r % 2 != 0 or 4 * r ** 3 > 256 * N)
        self.knn_after_front, self.join_before_devox = bool(knn_after_front), bool(join_before_devox)
        nb = -(-B // self.grid_chunks)
        self._chunks = [(b0, min(B, b0 + nb)) for b0 in range(0, B, nb)]
        f32, i32, dev = torch.float32, torch.int32, self.device
        with torch.cuda.device(dev):
            # inputs and host-visible outputs live in ONE device buffer each, mirrored by one pinned host buffer each: the
            # host-facing call moves them with a single copy per direction (five separate copies per step left ~10 % of the
            # PCIe link idle between transfers)
            ec = (2 if self.edge_echo else 1) * C

            def carve(total_shapes, device=None, pinned=False):
                offs, n = [], 0
                for shp in total_shapes:
                    offs.append(n)
                    cnt = 1
                    for d in shp:
                        cnt *= d
                    n += (cnt + 63) // 64 * 64                    # every view starts 256-byte aligned
                buf = torch.zeros(max(n, 1), dtype=f32).pin_memory() if pinned else torch.zeros(max(n, 1), dtype=f32, device=device)
                views = []
                for shp, o in zip(total_shapes, offs):
                    cnt = 1
                    for d in shp:
                        cnt *= d
                    views.append(buf[o:o + cnt].view(shp))
                return buf, views
            in_shapes = [(B, 6, N), (B, C, N)]
            out_shapes = [(B, 4, k, N), (B, C, N), (B, ec, N)]
            self._in_pack, (self.points, self.features) = carve(in_shapes, device=dev)
            self._out_pack, (self.ppf, self.devox, self.edge) = carve(out_shapes, device=dev)
            # branch A
            self.knn_dist = torch.empty((B, k, N), dtype=f32, device=dev)
            self.knn_idx = torch.empty((B, k, N), dtype=i32, device=dev)
            # branch B
            self.grid = torch.empty((B, C, r, r, r), dtype=f32, device=dev)
            self.ind = torch.empty((B, N), dtype=i32, device=dev)
            self.cnt = torch.empty((B, s), dtype=i32, device=dev)
            self.devox_inds = torch.empty((B, 8, N), dtype=i32, device=dev)
            self.devox_wgts = torch.empty((B, 8, N), dtype=f32, device=dev)
            self.xyz = torch.empty((B, 3, N), dtype=f32, device=dev)
            self.normals = torch.empty((B, 3, N), dtype=f32, device=dev)
            self._packed = torch.empty((B, N, 8), dtype=f32, device=dev)      # (x,y,z,nx,ny,nz,0,0) per point
            self.norm_coords = torch.empty((B, 3, N), dtype=f32, device=dev)
            self._vox_coords = torch.empty((B, 3, N), dtype=i32, device=dev)
            self._ws_bytes = _L.ri_voxelize_workspace_bytes(B, C, N, r)
            self._ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=dev)
            # both branches at the same stream priority: with the k-NN / PPF branch on a lower-priority stream (round 1) the
            # block scheduler served every pending CTA of the other batch's voxel branch first and the side branch ran in the
            # gaps only — 107.8 against 102.6 us per step with two batches in flight (sph_dg), 126.0 against 118.1 (cu_dg)
            self._side = torch.cuda.Stream(device=dev, priority=-1)
            self._main = torch.cuda.Stream(device=dev, priority=-1)
            self._devox_stream = torch.cuda.Stream(device=dev, priority=-1)
            # host staging (pinned)
            self._h_in_pack, (self.h_points, self.h_features) = carve(in_shapes, pinned=True)
            self._h_out_pack, (self.h_ppf, self.h_devox, self.h_edge) = carve(out_shapes, pinned=True)
        self._graph = None
        self._fused_front = self.grid_chunks == 1 and N <= 1024 and s % 4 == 0
        if not self.edge_echo and not self._fused_front:
            raise ValueError('edge_echo=False needs the fused prefix kernel (N <= 1024, r^3 % 4 == 0, one grid chunk)')
        # in-kernel mean (torch's reduction order, csrc/voxelize.cu): candidate only where that order is the one analysed
        # (128 <= N, N % 4 == 0); switched on by _verify_own_mean() at the first forward(), never silently
        self._own_mean = False
        # 3 B >= 16 outputs: below that torch widens its blocks (more lanes per output, another summation tree)
        self._own_mean_checked = not (fuse_mean and self._fused_front and N >= 128 and N % 4 == 0 and 3 * B >= 16)
        with torch.cuda.device(dev):
            self._mean_buf = torch.zeros((B, 3), dtype=f32, device=dev)

    # bytes moved per host-facing call
    @property
    def h2d_bytes(self):
        return self.h_points.numel() * 4 + self.h_features.numel() * 4

    @property
    def d2h_bytes(self):
        return (self.h_ppf.numel() + self.h_devox.numel() + self.h_edge.numel()) * 4

    # ------------------------------------------------------------------ the step, enqueued on current stream(s)
    def _knn(self):
        st = torch.cuda.current_stream().cuda_stream
        B, N, k = self.B, self.N, self.k
        _check(_L.ri_split_xyz_normals_f32(self.points.data_ptr(), B, N, self.xyz.data_ptr(), self.normals.data_ptr(),
                                           self._packed.data_ptr(), st), 'ri_split_xyz_normals')
        _check(_L.ri_knn_f32(self.xyz.data_ptr(), self.xyz.data_ptr(), B, 3, N, N, k,
                             self.knn_dist.data_ptr(), self.knn_idx.data_ptr(), st), 'ri_knn')

    def _ppf(self):
        st = torch.cuda.current_stream().cuda_stream
        _check(_L.ri_ppf_gather_packed_f32(self._packed.data_ptr(), self.knn_idx.data_ptr(), self.B, self.N, self.k,
                                           self.ppf.data_ptr(), st), 'ri_ppf_gather_packed')

    def _branch_a(self):
        self._knn()
        self._ppf()

    def _branch_b(self, join=None, fork=None):
        st = torch.cuda.current_stream().cuda_stream
        B, N, C, r = self.B, self.N, self.C, self.r
        # Coordinate prologue: bit-identical to the module shells (modules/voxelization.py) — asserted by tests/test_parity_gpu.py.
        sph = self.voxel_shape == 'spherical'
        shape = 2 if sph else (1 if self.normalize else 0)
        fused = self.grid_chunks == 1 and N <= 1024 and (r ** 3) % 4 == 0
        own_mean = fused and self._own_mean
        # The per-cloud mean: torch's own reduction (its summation order defines the bits the binning must see), or — on the
        # fused path, once _verify_own_mean() has seen it reproduce torch's bits for this shape — the same order inside the kernel
        mean = self._mean_buf if own_mean else self.points[:, :3, :].mean(2)
        if fused:
            # prologue + prepare + means/edge in one launch, then the grid writer
            _check(_L.ri_vox_front_f32(self.points.data_ptr(), 6, mean.data_ptr(), self.features.data_ptr(), B, C, N, r,
                                       shape, float(self.eps), self._front_mode(own_mean), self.norm_coords.data_ptr(),
                                       self._vox_coords.data_ptr(), self.ind.data_ptr(), self.edge.data_ptr(),
                                       self._ws.data_ptr(), self._ws_bytes, st), 'ri_vox_front')
            if fork is not None:
                # The k-NN starts once the latency-bound prefix has had the machine to itself: started at t = 0 the
                # long k-NN CTAs take the SMs first and stretch the prefix from 33 us to 85 us (tools/timeline.py);
                # measured step 204 us this way, 216-224 us with both branches released together.
                self._side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._side):
                    self._knn()
                    if join is None:
                        self._ppf()            # one carveout family: the side branch simply runs to its end
            _check(_L.ri_voxelize_fill_f32(B, C, N, r, 0, B, self.grid.data_ptr(), self.cnt.data_ptr(),
                                           self._ws.data_ptr(), self._ws_bytes, st), 'ri_voxelize_fill')
            if join is not None:
                torch.cuda.current_stream().wait_stream(join)      # see below: the devoxelizer runs after the k-NN
            if fork is not None and join is not None:
                # PPF (no shared memory, fp64-ALU bound) next to the devoxelizer (gather bound): both run with the
                # default max-L1 carveout, so they share the SMs once the max-shared kernels are gone
                self._side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._side):
                    self._ppf()
            self._devox(0, B, st)
            return
        _check(_L.ri_vox_prologue_f32(self.points.data_ptr(), 6, mean.data_ptr(), B, N, r, shape, float(self.eps),
                                      self.NORM_MODE, None, None, self.norm_coords.data_ptr(),
                                      self._vox_coords.data_ptr(), st), 'ri_vox_prologue')
        nc = self.norm_coords
        if self.grid_chunks == 1:
            if sph:
                _check(_L.ri_sph_voxelize_edge_f32(self.features.data_ptr(), nc.data_ptr(), B, C, N, r,
                                                   self.grid.data_ptr(), self.ind.data_ptr(), self.cnt.data_ptr(),
                                                   self.edge.data_ptr(), self._ws.data_ptr(), self._ws_bytes, st),
                       'ri_sph_voxelize_edge')
            else:
                _check(_L.ri_cube_voxelize_edge_f32(self.features.data_ptr(), self._vox_coords.data_ptr(), B, C, N, r,
                                                    self.grid.data_ptr(), self.ind.data_ptr(), self.cnt.data_ptr(),
                                                    self.edge.data_ptr(), self._ws.data_ptr(), self._ws_bytes, st),
                       'ri_cube_voxelize_edge')
            if join is not None:
                # the devoxelizer's gathers live off a large L1; CTAs of the k-NN branch (max-shared carveout) on the
                # same SMs would force the small-L1 split on it (measured 54 us -> 300 us), so it starts after them
                torch.cuda.current_stream().wait_stream(join)
            self._devox(0, B, st)
            return
        if sph:
            _check(_L.ri_sph_voxelize_prepare_f32(nc.data_ptr(), B, C, N, r, self.ind.data_ptr(),
                                                  self._ws.data_ptr(), self._ws_bytes, st), 'ri_sph_voxelize_prepare')
        else:
            _check(_L.ri_cube_voxelize_prepare_f32(self._vox_coords.data_ptr(), B, C, N, r, self.ind.data_ptr(),
                                                   self._ws.data_ptr(), self._ws_bytes, st), 'ri_cube_voxelize_prepare')
        main = torch.cuda.current_stream()
        for (b0, b1) in self._chunks:
            _check(_L.ri_voxelize_means_f32(self.features.data_ptr(), B, C, N, r, b0, b1, self.edge.data_ptr(),
                                            self._ws.data_ptr(), self._ws_bytes, st), 'ri_voxelize_means')
            _check(_L.ri_voxelize_fill_f32(B, C, N, r, b0, b1, self.grid.data_ptr(), self.cnt.data_ptr(),
                                           self._ws.data_ptr(), self._ws_bytes, st), 'ri_voxelize_fill')
            if self.devox_side_stream:
                self._devox_stream.wait_stream(main)
                with torch.cuda.stream(self._devox_stream):
                    self._devox(b0, b1, self._devox_stream.cuda_stream)
            else:
                self._devox(b0, b1, st)
        if self.devox_side_stream:
            main.wait_stream(self._devox_stream)

    def _front_mode(self, own_mean):
        """norm_mode argument of ri_vox_front_f32: radius association | in-kernel mean | relative-half-only edge output."""
        return self.NORM_MODE | (0x100 if own_mean else 0) | (0 if self.edge_echo else 0x200)

    def _devox(self, b0, b1, st):
        """Devoxelize the clouds [b0, b1) of the batch (pointer offsets into the whole-batch arrays)."""
        N, C, r = self.N, self.C, self.r
        s = r ** 3
        nb = b1 - b0
        nc, g = self.norm_coords.data_ptr() + b0 * 3 * N * 4, self.grid.data_ptr() + b0 * C * s * 4
        outs = self.devox.data_ptr() + b0 * C * N * 4
        inds, wgts = self.devox_inds.data_ptr() + b0 * 8 * N * 4, self.devox_wgts.data_ptr() + b0 * 8 * N * 4
        if self.voxel_shape == 'spherical':
            _check(_L.ri_sph_trilinear_devox_f32(nc, g, self.ind.data_ptr() + b0 * N * 4, nb, C, N, r, outs, inds, wgts, st),
                   'ri_sph_trilinear_devox')
        else:
            _check(_L.ri_trilinear_devox_f32(nc, g, nb, C, N, r, outs, inds, wgts, st), 'ri_trilinear_devox')

    def _step(self):
        if self.overlap:
            cur = torch.cuda.current_stream()
            self._main.wait_stream(cur)
            self._side.wait_stream(cur)
            if self.knn_after_front and self._fused_front:
                with torch.cuda.stream(self._main):
                    self._branch_b(join=self._side if self.join_before_devox else None, fork=self._branch_a)
            else:
                with torch.cuda.stream(self._side):
                    self._branch_a()
                with torch.cuda.stream(self._main):
                    self._branch_b(join=self._side if self.join_before_devox else None)
            cur.wait_stream(self._main)
            cur.wait_stream(self._side)
        else:
            self._branch_a()
            self._branch_b()

    def _verify_own_mean(self):
        """Run the fused prefix once with the in-kernel mean on a seeded random probe batch of this engine's shape and compare
        the means it wrote with torch's `probe[:, :3].mean(2)` bit for bit; only then is the torch reduction dropped from the
        step (a torch build with another reduction heuristic simply keeps the torch kernel)."""
        self._own_mean_checked = True
        B, N, C, r = self.B, self.N, self.C, self.r
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            shape = 2 if self.voxel_shape == 'spherical' else (1 if self.normalize else 0)
            g = torch.Generator(device=self.device); g.manual_seed(20261018)
            ok = True
            for stride in (7, 3, 11):                      # three probes: 9 B sums have to come out bit-identical
                probe = torch.randn((B, 6, N), dtype=torch.float32, device=self.device, generator=g) * 3.0 + 0.7
                probe[:, :3, ::stride] *= 1e3              # mixed magnitudes: any other summation order changes low bits
                rc = _L.ri_vox_front_f32(probe.data_ptr(), 6, self._mean_buf.data_ptr(), self.features.data_ptr(), B, C, N, r,
                                         shape, float(self.eps), self._front_mode(True), self.norm_coords.data_ptr(),
                                         self._vox_coords.data_ptr(), self.ind.data_ptr(), self.edge.data_ptr(),
                                         self._ws.data_ptr(), self._ws_bytes, st)
                if rc != 0:
                    return
                want = probe[:, :3, :].mean(2)
                ok = ok and bool(torch.equal(want, self._mean_buf)) and bool(torch.isfinite(want).all())
            self._own_mean = ok

    def _capture(self):
        if not self._own_mean_checked:
            self._verify_own_mean()
        with torch.cuda.device(self.device):
            warm = torch.cuda.Stream(device=self.device)
            warm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(warm):
                for _ in range(2):
                    self._step()
            torch.cuda.current_stream().wait_stream(warm)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step()
            self._graph = g

    def forward(self):
        """One step on the device-resident `points` / `features` buffers (asynchronous)."""
        with torch.cuda.device(self.device):
            if not self.use_graph:
                if not self._own_mean_checked:
                    self._verify_own_mean()
                self._step()
                return
            if self._graph is None:
                self._capture()
            self._graph.replay()

    def load(self, points, features):
        """Copy a batch ([B,6,N], [B,C,N]; numpy or torch, host or device) into the device input buffers."""
        self.points.copy_(torch.as_tensor(points), non_blocking=True)
        self.features.copy_(torch.as_tensor(features), non_blocking=True)

    def __call__(self, points_host, features_host):
        """Host-facing call: host arrays in, pinned host tensors out {ppf, devox, edge} (synchronous)."""
        with torch.cuda.device(self.device):
            self.h_points.copy_(torch.as_tensor(points_host))
            self.h_features.copy_(torch.as_tensor(features_host))
            return self.run_staged()

    def run_staged(self):
        """H2D from the pinned staging buffers -> step -> D2H into pinned buffers, then wait."""
        with torch.cuda.device(self.device):
            self._in_pack.copy_(self._h_in_pack, non_blocking=True)
            self.forward()
            self._h_out_pack.copy_(self._out_pack, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return {'ppf': self.h_ppf, 'devox': self.h_devox, self._edge_key: self.h_edge}

    @property
    def _edge_key(self):
        return 'edge' if self.edge_echo else 'edge_rel'

    # algorithmic (compulsory) bytes of one step, SURVEY.md §8(d) formulas, fused KNN->PPF form
    def algorithmic_bytes(self):
        B, N, C, k, r = self.B, self.N, self.C, self.k, self.r
        s = r ** 3
        knn_ppf = 24 * N + 16 * k * N
        vox = 12 * N + 4 * C * N + 4 * N + 4 * s + 4 * C * s
        devox = 12 * N + 4 * N + min(32 * C * N, 4 * C * s) + 4 * C * N + 64 * N
        edge = 4 * N + 4 * C * N + 4 * C * N + (8 if self.edge_echo else 4) * C * N
        return {'knn_ppf': B * knn_ppf, 'voxelize': B * vox, 'devox': B * devox, 'edge': B * edge,
                'total': B * (knn_ppf + vox + devox + edge)}


class FrontEndLanes:
    """Keeps several steps in flight on the device: step i runs engine i % len(engines) on launch stream i % lanes.
    Consecutive steps are independent batches, so the latency-bound prefix of step i+1 (mean, cell sort, cell means) and its
    ALU-bound k-NN run under the HBM-bound grid write / devoxelize of step i.  An engine is never replayed before its own
    previous step has finished (per-engine event).

        lanes = FrontEndLanes(engines, lanes=2)     # engines built with fuse_mean=False: best throughput in flight
        lanes.begin()                 # the launch streams pick up after the current stream
        for i in range(K): lanes.forward(i)
        lanes.end()                   # the current stream waits for every step in flight
    """

    def __init__(self, engines, lanes=2):
        self.engines = list(engines)
        self.device = self.engines[0].device
        self.lanes = int(lanes)
        with torch.cuda.device(self.device):
            self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)]
            self._done = [None] * len(self.engines)
            for fe in self.engines:                # graphs are captured up front: a capture cannot overlap other work
                fe.forward()
            torch.cuda.synchronize(self.device)

    def begin(self):
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)

    def forward(self, i):
        e = i % len(self.engines)
        lane = self.streams[i % self.lanes]
        if self._done[e] is not None:
            lane.wait_event(self._done[e])
        with torch.cuda.stream(lane):
            self.engines[e].forward()
            ev = self._done[e] if self._done[e] is not None else torch.cuda.Event()
            ev.record(lane)
            self._done[e] = ev

    def end(self):
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)


class FrontEndPipeline:
    """Host-facing streaming form of the engine: `depth` FrontEnd slots, each with its own pinned host buffers and device
    buffers, driven by three streams so that consecutive calls overlap — while slot i computes, slot i+1's inputs go up
    (H2D) and slot i-1's outputs come down (D2H; PCIe is full duplex).  Every call still moves its own inputs host->device
    and its own outputs device->host; only the waiting is taken out.

        pipe = FrontEndPipeline(B, N, C, depth=3, ...)
        s = pipe.acquire()                  # next slot, its previous results have been handed back
        pipe.slot(s).h_points[...] = ...    # fill the pinned inputs
        pipe.submit(s)                      # H2D -> step -> D2H, asynchronous
        out = pipe.result(s)                # {'ppf', 'devox', 'edge' | 'edge_rel'} pinned host tensors (waits for that slot only)
    """

    def __init__(self, B, N, C, depth=3, device='cuda', **kw):
        self.depth = int(depth)
        self.device = torch.device(device)
        self.slots = [FrontEnd(B, N, C, device=device, **kw) for _ in range(self.depth)]
        with torch.cuda.device(self.device):
            self._up = torch.cuda.Stream(device=self.device)
            self._run = torch.cuda.Stream(device=self.device)
            self._down = torch.cuda.Stream(device=self.device)
            mk = lambda: [torch.cuda.Event() for _ in range(self.depth)]
            self._ev_up, self._ev_run, self._ev_down = mk(), mk(), mk()
            self._busy = [False] * self.depth
            self._next = 0
            for fe in self.slots:                 # capture the graphs up front (capture cannot overlap other work)
                with torch.cuda.stream(self._run):
                    fe.forward()
            torch.cuda.synchronize(self.device)

    def slot(self, s):
        return self.slots[s]

    def acquire(self):
        """Next slot in round-robin order; blocks until the results of its previous submission have landed."""
        s = self._next
        self._next = (s + 1) % self.depth
        if self._busy[s]:
            self._ev_down[s].synchronize()
            self._busy[s] = False
        return s

    def submit(self, s):
        fe = self.slots[s]
        with torch.cuda.device(self.device):
            # the slot's device inputs are free once its previous step has run; its device outputs once they were copied
            self._up.wait_event(self._ev_run[s])
            with torch.cuda.stream(self._up):
                fe._in_pack.copy_(fe._h_in_pack, non_blocking=True)
                self._ev_up[s].record(self._up)
            self._run.wait_event(self._ev_up[s])
            self._run.wait_event(self._ev_down[s])
            with torch.cuda.stream(self._run):
                fe.forward()
                self._ev_run[s].record(self._run)
            self._down.wait_event(self._ev_run[s])
            with torch.cuda.stream(self._down):
                fe._h_out_pack.copy_(fe._out_pack, non_blocking=True)
                self._ev_down[s].record(self._down)
        self._busy[s] = True

    def result(self, s):
        self._ev_down[s].synchronize()
        self._busy[s] = False
        fe = self.slots[s]
        return {'ppf': fe.h_ppf, 'devox': fe.h_devox, fe._edge_key: fe.h_edge}

    def drain(self):
        for s in range(self.depth):
            if self._busy[s]:
                self._ev_down[s].synchronize()
                self._busy[s] = False
