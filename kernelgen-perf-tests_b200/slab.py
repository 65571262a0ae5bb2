"""z-slab engine: one process per GPU, the grid cut into contiguous slabs along the slowest
dimension (z for the 3D tests, y for the 2D tests), ghost planes refreshed once per sweep.

Host-side logic only -- the sweeps are launched through the C ABI (capi.sweep).  Two ways to
refresh ghosts:

  halo="push"  the sweep kernel stores the planes a neighbour needs straight into the
               neighbour's memory (CUDA-IPC peer pointer, NVLink); ordering between ranks is a
               flag per neighbour handled INSIDE the sweep kernel (spin at kernel start, release
               store by the last CTA), the whole nt-loop is one C call (b200_slab_loop): no host
               round trip and no collective on the data path;
  halo="nccl"  torch.distributed batched isend/irecv of the boundary planes after each sweep
               (also what the CPU/gloo tests use).

torch provides device memory, streams, events and torch.distributed; nothing else.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


class SlabLayout:
    """Pure arithmetic: which planes a rank owns / stores / exchanges."""

    def __init__(self, info: dict, n_global: int, world: int, rank: int):
        self.info, self.n, self.world, self.rank = info, n_global, world, rank
        self.split_dim = 2 if info["ndims"] == 3 else 1
        glo, ghi = info["zghost_lo"], info["zghost_hi"]
        self.own_lo = n_global * rank // world
        self.own_hi = n_global * (rank + 1) // world
        self.mem_lo = max(0, self.own_lo - glo)
        self.mem_hi = min(n_global, self.own_hi + ghi)
        if world > 1 and self.own_hi - self.own_lo < glo + ghi + 1:
            raise ValueError(f"extent {n_global} too small for {world} slabs")
        self.lo_ghost = self.own_lo - self.mem_lo        # planes received from rank-1
        self.hi_ghost = self.mem_hi - self.own_hi        # planes received from rank+1
        # what the neighbours need from us
        self.send_lo_cnt = ghi if rank > 0 else 0          # our lowest owned planes -> rank-1's upper ghosts
        self.send_hi_cnt = glo if rank < world - 1 else 0  # our highest owned planes -> rank+1's lower ghosts

    @property
    def mem_n(self):
        return self.mem_hi - self.mem_lo

    def out_range(self):
        """Owned planes inside the global interior, in LOCAL coordinates (half-open)."""
        lo = max(self.own_lo, self.info["lo"][self.split_dim])
        hi = min(self.own_hi, self.n - self.info["hi"][self.split_dim])
        return lo - self.mem_lo, max(hi, lo) - self.mem_lo

    # local plane ranges
    def send_lo(self):
        a = self.own_lo - self.mem_lo
        return a, a + self.send_lo_cnt

    def send_hi(self):
        b = self.own_hi - self.mem_lo
        return b - self.send_hi_cnt, b

    def recv_lo(self):
        return 0, self.lo_ghost

    def recv_hi(self):
        b = self.own_hi - self.mem_lo
        return b, b + self.hi_ghost


def exchange_halos(dist, layout: SlabLayout, t):
    """Refresh the ghost planes of tensor `t` (leading dimension = split dimension) with the
    neighbours' boundary planes: torch.distributed point-to-point, any backend."""
    ops = []
    r = layout.rank
    if layout.send_lo_cnt:
        a, b = layout.send_lo()
        ops.append(dist.P2POp(dist.isend, t[a:b], r - 1))
    if layout.send_hi_cnt:
        a, b = layout.send_hi()
        ops.append(dist.P2POp(dist.isend, t[a:b], r + 1))
    if layout.lo_ghost:
        a, b = layout.recv_lo()
        ops.append(dist.P2POp(dist.irecv, t[a:b], r - 1))
    if layout.hi_ghost:
        a, b = layout.recv_hi()
        ops.append(dist.P2POp(dist.irecv, t[a:b], r + 1))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class _DevMem:
    """cudaMalloc'ed memory from the C ABI (IPC-exportable), viewable as a torch tensor."""

    def __init__(self, pkg, nelem, np_dtype):
        self.pkg, self.nelem, self.dtype = pkg, nelem, np.dtype(np_dtype)
        self.ptr = pkg.capi.device_alloc(nelem * self.dtype.itemsize)
        self.__cuda_array_interface__ = {"shape": (nelem,), "typestr": self.dtype.str, "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def tensor(self, torch):
        return torch.as_tensor(self, device="cuda")

    def free(self):
        if self.ptr:
            self.pkg.capi.device_free(self.ptr)
            self.ptr = 0


class SlabEngine:
    """One rank's slab of one test: buffers, rotation, sweeps, ghost refresh."""

    def __init__(self, pkg, test, real, nx, ny, ns, scalars, world=1, rank=0, dist=None, halo="push", seed=0):
        import torch
        self.torch, self.pkg, self.dist = torch, pkg, dist
        self.test, self.real, self.scalars = test, real, list(scalars)
        self.info = info = pkg.test_info(test)
        self.world, self.rank, self.halo = world, rank, halo
        self.nx, self.ny, self.ns = nx, ny, (ns if info["ndims"] == 3 else 1)
        self.np_dtype = np.float32 if real == "float" else np.float64
        per_rank = self.ns if info["ndims"] == 3 else ny
        self.layout = L = SlabLayout(info, per_rank * world, world, rank)
        self.exchange = world > 1 and info["exchange_slot"] >= 0
        self.unit = nx * ny if info["ndims"] == 3 else nx            # elements per plane / row
        if test == "matvec":
            self.unit = nx
        if test == "matmul":
            self.unit = nx          # columns of C; B's columns hold ny elements, A is replicated
        self.stream = torch.cuda.current_stream()
        # buffers: slab incl. ghosts, from the C ABI allocator (IPC-exportable)
        self.mem, self.t = [], []
        g = torch.Generator(device="cuda")
        g.manual_seed(seed)
        for q in range(info["narrays"]):
            n = self._slot_len(q)
            m = _DevMem(pkg, n, self.np_dtype)
            t = m.tensor(torch)
            t.uniform_(-1.0, 1.0, generator=g)           # in place: no temporaries next to 30+ GB slabs
            self.mem.append(m)
            self.t.append(t)
        # temporal blocking (1 GPU, tests with a fused two-sweep kernel): a third buffer holding w0's shell
        self.scratch = self.scratch_t = None
        if world == 1 and pkg.capi.sweep2_profitable(test, nx):
            self.scratch = _DevMem(pkg, self._slot_len(0), self.np_dtype)
            self.scratch_t = self.scratch.tensor(torch)
            self.scratch_t.copy_(self.t[0])
        self.scratch_shell = 0          # index into self.mem whose boundary shell the scratch buffer carries
        self.idxs = [0, 1, 2]
        self.sweeps_done = 0
        self.peer = {}
        if self.exchange:
            # consistent ghosts to start from (all rotating buffers)
            for q in range(info["rotation"]):
                exchange_halos(dist, L, self.t[q].view(L.mem_n, -1))
            if halo == "push":
                self._setup_push()
        torch.cuda.synchronize()

    # -- sizes -------------------------------------------------------------------------------
    def _slot_len(self, q):
        if self.test == "matvec":
            return [self.nx * self.layout.mem_n, self.nx, self.layout.mem_n][q]
        if self.test == "matmul":
            return [self.nx * self.ny, self.ny * self.layout.mem_n, self.nx * self.layout.mem_n][q]
        return self.unit * self.layout.mem_n

    def local_dims(self):
        L = self.layout
        if self.info["ndims"] == 3:
            return self.nx, self.ny, L.mem_n
        return self.nx, L.mem_n, 1

    def local_interior_points(self):
        a, b = self.layout.out_range()
        i = self.info
        if self.test == "matvec":
            return self.nx * (b - a)
        if self.test == "matmul":
            return self.nx * self.ny * (b - a)
        ex = self.nx - i["lo"][0] - i["hi"][0]
        ey = (self.ny - i["lo"][1] - i["hi"][1]) if i["ndims"] == 3 else 1
        return max(ex, 0) * max(ey, 0) * max(b - a, 0)

    def global_interior_points(self):
        L = self.layout
        if self.info["ndims"] == 3:
            return self.pkg.interior_points(self.test, self.nx, self.ny, L.n)
        return self.pkg.interior_points(self.test, self.nx, L.n, 1)

    # -- push mode plumbing --------------------------------------------------------------------
    def _setup_push(self):
        torch, dist, capi = self.torch, self.dist, self.pkg.capi
        nrot = self.info["rotation"]
        self.flags = _DevMem(self.pkg, 8, np.int64)             # [0]: written by rank-1, [1]: by rank+1
        self.flags.tensor(torch).zero_()
        torch.cuda.synchronize()
        mine = {"bufs": [capi.ipc_export(self.mem[q].ptr) for q in range(nrot)],
                "flags": capi.ipc_export(self.flags.ptr)}
        allh = [None] * self.world
        dist.all_gather_object(allh, mine)
        for nb in (self.rank - 1, self.rank + 1):
            if 0 <= nb < self.world:
                self.peer[nb] = {"bufs": [capi.ipc_import(h) for h in allh[nb]["bufs"]],
                                 "flags": capi.ipc_import(allh[nb]["flags"])}
        dist.barrier()

    def _push_desc(self, out_slot):
        """Peer pointers + plane ranges for the fused halo push of this sweep's output array."""
        L, push = self.layout, {}
        esz = np.dtype(self.np_dtype).itemsize
        peers = getattr(self, "_peer_layouts", None)
        if peers is None:
            peers = self._peer_layouts = {nb: SlabLayout(self.info, L.n, self.world, nb) for nb in self.peer}
        if self.rank - 1 in self.peer:
            n = peers[self.rank - 1]
            a, _ = L.send_lo()
            push["lo"] = (self.peer[self.rank - 1]["bufs"][out_slot], a, n.own_hi - n.mem_lo, L.send_lo_cnt)
        if self.rank + 1 in self.peer:
            a, _ = L.send_hi()
            push["hi"] = (self.peer[self.rank + 1]["bufs"][out_slot], a, 0, L.send_hi_cnt)
        return push

    # -- the hot loop --------------------------------------------------------------------------
    def run(self, niters: int):
        """`niters` sweeps with the reference driver's buffer rotation (laplacian.c:287-301)."""
        capi, info, L = self.pkg.capi, self.info, self.layout
        nx, ny, ns = self.local_dims()
        rot = info["rotation"]
        out_pos = 2 if rot == 3 else 1
        stream = self.stream.cuda_stream
        a, b = L.out_range()
        out_range = (a, b) if (self.world > 1 or self.test in ("vecadd", "sincos", "matvec", "matmul")) else None
        if not self.exchange:
            # no ghost refresh between sweeps: the whole nt-loop is enqueued by one C call
            if b > a:
                ptrs = [m.ptr for m in self.mem]
                for q in range(rot):
                    ptrs[q] = self.mem[self.idxs[q]].ptr
                if self.scratch is not None and out_range is None:
                    n2 = niters
                    if self.idxs[0] != self.scratch_shell and niters >= 1:
                        # after an odd number of sweeps the w0 role is held by the OTHER buffer, whose shell the scratch
                        # does not carry: one single sweep first realigns the roles (b200_run does the same)
                        capi.sweep_loop(self.test, self.real, nx, ny, ns, self.scalars, ptrs, 1, stream=stream)
                        ptrs[0], ptrs[1] = ptrs[1], ptrs[0]
                        n2 = niters - 1
                    _, scr = capi.sweep_loop2(self.test, self.real, nx, ny, ns, self.scalars, ptrs, self.scratch.ptr,
                                              n2, stream=stream)
                    if scr != self.scratch.ptr:      # an odd number of fused passes: old w0 buffer <-> scratch
                        q = next(i for i, m in enumerate(self.mem) if m.ptr == scr)
                        self.mem[q], self.scratch = self.scratch, self.mem[q]
                        self.t[q], self.scratch_t = self.scratch_t, self.t[q]
                else:
                    capi.sweep_loop(self.test, self.real, nx, ny, ns, self.scalars, ptrs, niters, stream=stream,
                                    out_range=out_range)
            for _ in range(niters):
                if rot == 2:
                    self.idxs[0], self.idxs[1] = self.idxs[1], self.idxs[0]
                elif rot == 3:
                    self.idxs = [self.idxs[1], self.idxs[2], self.idxs[0]]
            self.sweeps_done += niters
            return
        if self.halo == "push" and b > a:
            # fused halo push, neighbour ordering inside the sweep kernel, the whole loop in one C call
            na = len(self.mem)
            ptrs = [m.ptr for m in self.mem]
            plo, phi = [0] * na, [0] * na
            for q in range(rot):
                ptrs[q] = self.mem[self.idxs[q]].ptr
                if self.rank - 1 in self.peer:
                    plo[q] = self.peer[self.rank - 1]["bufs"][self.idxs[q]]
                if self.rank + 1 in self.peer:
                    phi[q] = self.peer[self.rank + 1]["bufs"][self.idxs[q]]
            peers = getattr(self, "_peer_layouts", None)
            if peers is None:
                peers = self._peer_layouts = {nb: SlabLayout(self.info, L.n, self.world, nb) for nb in self.peer}
            push_lo, push_hi = (0, 0, 0), (0, 0, 0)
            wait, sig = [0, 0], [0, 0]
            if self.rank - 1 in self.peer:
                n = peers[self.rank - 1]
                push_lo = (L.send_lo()[0], n.own_hi - n.mem_lo, L.send_lo_cnt)
                wait[0] = self.flags.ptr
                sig[0] = self.peer[self.rank - 1]["flags"] + 8       # we are its upper neighbour
            if self.rank + 1 in self.peer:
                push_hi = (L.send_hi()[0], 0, L.send_hi_cnt)
                wait[1] = self.flags.ptr + 8
                sig[1] = self.peer[self.rank + 1]["flags"]           # we are its lower neighbour
            capi.slab_loop(self.test, self.real, nx, ny, ns, self.scalars, ptrs, plo, phi, niters,
                           self.sweeps_done, (a, b), push_lo, push_hi, wait, sig, stream=stream)
            for _ in range(niters):
                if rot == 2:
                    self.idxs[0], self.idxs[1] = self.idxs[1], self.idxs[0]
                elif rot == 3:
                    self.idxs = [self.idxs[1], self.idxs[2], self.idxs[0]]
            self.sweeps_done += niters
            return
        for _ in range(niters):
            ptrs = [m.ptr for m in self.mem]
            for q in range(rot):
                ptrs[q] = self.mem[self.idxs[q]].ptr
            push = None
            if self.exchange and self.halo == "push":
                it = self.sweeps_done
                if it > 0:
                    # ghosts of this sweep's input were pushed during the neighbours' previous sweep
                    if self.rank - 1 in self.peer:
                        capi.wait_flag(self.flags.ptr, it, stream)
                    if self.rank + 1 in self.peer:
                        capi.wait_flag(self.flags.ptr + 8, it, stream)
                push = self._push_desc(self.idxs[out_pos])
            if b > a:
                capi.sweep(self.test, self.real, nx, ny, ns, self.scalars, ptrs, stream=stream,
                           out_range=out_range, push=push)
            if self.exchange:
                if self.halo == "push":
                    it = self.sweeps_done + 1
                    if self.rank - 1 in self.peer:      # we are rank-1's upper neighbour -> its flag[1]
                        capi.signal_flag(self.peer[self.rank - 1]["flags"] + 8, it, stream)
                    if self.rank + 1 in self.peer:      # we are rank+1's lower neighbour -> its flag[0]
                        capi.signal_flag(self.peer[self.rank + 1]["flags"], it, stream)
                else:
                    exchange_halos(self.dist, L, self.t[self.idxs[out_pos]].view(L.mem_n, -1))
            if rot == 2:
                self.idxs[0], self.idxs[1] = self.idxs[1], self.idxs[0]
            elif rot == 3:
                self.idxs = [self.idxs[1], self.idxs[2], self.idxs[0]]
            self.sweeps_done += 1

    # -- deterministic contents (parity checks at any size) -------------------------------------
    def fill_pattern(self, seed: int = 0):
        """Every array becomes a fixed function of the GLOBAL element index (an integer hash mapped to [-1, 1)), so any
        decomposition of the same global grid holds the same values at the same points, ghosts included -- no RNG state,
        no host array, no exchange.  Stencil tests only (uniform slot layout)."""
        torch, L = self.torch, self.layout
        assert self.test not in ("matvec", "matmul")
        chunk = 1 << 26
        for q, t in enumerate(self.t):
            base = L.mem_lo * self.unit
            for a in range(0, t.numel(), chunk):
                b = min(t.numel(), a + chunk)
                e = torch.arange(base + a, base + b, device=t.device, dtype=torch.int64)
                h = (e * 2654435761 + (q + 1) * 40503 + seed * 69069) & 0x7FFFFFFF
                h = (h ^ (h >> 13)) * 1274126177 & 0x7FFFFFFF
                t[a:b] = (h.to(torch.float64) * (2.0 / 2147483648.0) - 1.0).to(t.dtype)
        if self.scratch_t is not None:
            self.scratch_t.copy_(self.t[self.idxs[0]])
            self.scratch_shell = self.idxs[0]
        torch.cuda.synchronize()

    # -- scatter / gather of a global host array (tests) ---------------------------------------
    def set_global(self, q: int, full: np.ndarray):
        """Load this rank's slab (owned planes + ghosts) of slot q from the GLOBAL array."""
        L, torch = self.layout, self.torch
        if self.test == "matvec":
            part = {0: full[L.mem_lo * self.nx:L.mem_hi * self.nx], 1: full, 2: full[L.mem_lo:L.mem_hi]}[q]
        elif self.test == "matmul":
            part = {0: full, 1: full[L.mem_lo * self.ny:L.mem_hi * self.ny], 2: full[L.mem_lo * self.nx:L.mem_hi * self.nx]}[q]
        else:
            part = full[L.mem_lo * self.unit:L.mem_hi * self.unit]
        self.t[q].copy_(torch.from_numpy(np.ascontiguousarray(part)).to(self.t[q].device))
        if self.scratch_t is not None and q == self.idxs[0]:
            self.scratch_t.copy_(self.t[q])      # the fused path's third buffer carries the shell of the w0-role array
            self.scratch_shell = q

    def get_owned(self, q: int) -> np.ndarray:
        """This rank's OWNED planes of slot q (no ghosts), as a host array."""
        L = self.layout
        if self.test == "matvec":
            if q == 1:
                return self.t[q].cpu().numpy()
            u = self.nx if q == 0 else 1
        elif self.test == "matmul":
            if q == 0:
                return self.t[q].cpu().numpy()
            u = self.ny if q == 1 else self.nx
        else:
            u = self.unit
        a, b = (L.own_lo - L.mem_lo) * u, (L.own_hi - L.mem_lo) * u
        return self.t[q][a:b].cpu().numpy()

    def rewind(self):
        """A fresh job on the same buffers: slot q is the reference's array q again (the flag counters keep running)."""
        self.idxs = [0, 1, 2]

    def load_host(self, host, shell_only_dead=True):
        """Enqueue (on the engine's stream) the upload of this rank's slab of every array from pinned host tensors, the
        way the C drivers load: arrays whose interior the first sweep overwrites (b200_slot_interior_dead) send only their
        boundary shell (b200_load_shell_slab).  Before anything is overwritten the stream waits until the neighbours'
        previous job has stopped pushing into these buffers (their flags have reached this rank's sweep count)."""
        capi, L = self.pkg.capi, self.layout
        stream = self.stream.cuda_stream
        if self.exchange and self.halo == "push" and self.sweeps_done > 0:
            if self.rank - 1 in self.peer:
                capi.wait_flag(self.flags.ptr, self.sweeps_done, stream)
            if self.rank + 1 in self.peer:
                capi.wait_flag(self.flags.ptr + 8, self.sweeps_done, stream)
        nbytes = 0
        nx, ny, _ = self.local_dims()
        with self.torch.cuda.stream(self.stream):
            for q, h in enumerate(host):
                dead = shell_only_dead and capi.slot_interior_dead(self.test, q) and self.test not in ("matvec", "matmul")
                if dead:
                    capi.load_shell_slab(self.test, self.real, self.nx, self.ny, L.n, L.mem_lo, L.mem_hi, self.mem[q].ptr,
                                         h.data_ptr(), stream)
                    i = self.info
                    inner = max(self.nx - i["lo"][0] - i["hi"][0], 0) * (max(self.ny - i["lo"][1] - i["hi"][1], 0) if i["ndims"] == 3 else 1)
                    a, b = L.out_range()
                    nbytes += (h.numel() - inner * max(b - a, 0)) * h.element_size() if i["lo"][0] else 0
                else:
                    self.t[q].copy_(h, non_blocking=True)
                    nbytes += h.numel() * h.element_size()
        return nbytes

    def result_slot(self):
        if self.info["rotation"]:
            return self.idxs[1]
        return {"divergence": 0, "gradient": 1}.get(self.test, 2)

    # -- end to end on host buffers ---------------------------------------------------------------
    def e2e(self, niters, steps, barrier, dist=None):
        """Per step: H2D of every array from pinned host memory, niters sweeps, D2H of the result.
        N=1 goes through the driver-phase C ABI (b200_load / b200_run / b200_save); N>1 copies each
        rank's slab with cudaMemcpyAsync on the sweep stream."""
        import time
        torch, capi = self.torch, self.pkg.capi
        esz = np.dtype(self.np_dtype).itemsize
        nbytes_in = sum(self._slot_len(q) for q in range(self.info["narrays"])) * esz
        if self.world == 1:
            nx, ny, ns = self.local_dims()
            # Two contexts in asynchronous mode (b200_set_async), driven alternately by this one host thread: step i's
            # device->host copy and sweeps overlap step i+1's host->device copy (PCIe is full duplex, the copies sit
            # on each context's own stream).  Every step still uploads its own inputs from its own pinned buffers and
            # downloads its own result inside the timed region.  B200_E2E_PIPELINE=0: one context, phase by phase.
            import os
            depth = 2 if os.environ.get("B200_E2E_PIPELINE", "1") != "0" else 1
            rng = np.random.default_rng(7)
            lanes = []
            for _ in range(depth):
                host = [capi.PinnedBuffer(self._slot_len(q), self.np_dtype) for q in range(self.info["narrays"])]
                for h in host:
                    h.array[:] = rng.uniform(-1, 1, h.array.size).astype(self.np_dtype)
                ctx = capi.Context(1)
                ctx.plan(self.test, self.real, nx, ny, ns if self.info["ndims"] == 3 else 1, self.scalars)
                ctx.alloc()
                ctx.set_async(depth > 1)
                lanes.append((ctx, host))
            host = lanes[0][1]

            # what the C drivers do: output buffers (interior overwritten by the first sweep before anything
            # reads it) only send their boundary shell
            dead = [lanes[0][0].interior_dead(q) for q in range(len(host))]
            interior = self.pkg.interior_points(self.test, nx, ny, ns) if self.test not in ("matvec", "matmul") else 0
            nbytes_in = sum((h.array.size - interior if d and self.info["lo"][0] else (0 if d else h.array.size))
                            for h, d in zip(host, dead)) * esz

            def step(i):
                ctx, hb = lanes[i % depth]
                if depth > 1:
                    ctx.sync()                    # this lane's previous step has left its host buffers
                ctx.rewind()                      # a fresh job: slot q is the reference's array q again
                for q, h in enumerate(hb):
                    if dead[q]:
                        ctx.load_array_shell(q, h.array)
                    else:
                        ctx.load_array(q, h.array)
                ctx.run(niters)
                slot = ctx.result_slot()
                ctx.save_array(slot, hb[slot].array)
                return slot

            def drain():
                for ctx, _ in lanes:
                    ctx.sync()
                torch.cuda.synchronize()

            for i in range(depth):
                step(i)
            drain()
            t0 = time.perf_counter()
            for i in range(steps):
                slot = step(i)
            drain()
            dt = (time.perf_counter() - t0) / steps
            nbytes_out = host[slot].array.nbytes
            for ctx, hb in lanes:
                ctx.set_async(False)
                ctx.free()
                ctx.destroy()
                for h in hb:
                    h.free()
            api = "b200_load/b200_load_shell/b200_run/b200_save (C ABI, pinned host buffers)"
            if depth > 1:
                api += "; two contexts in b200_set_async mode, alternating (copy/compute overlap across steps)"
            return {"seconds_per_step": dt, "h2d_bytes_per_step": int(nbytes_in), "d2h_bytes_per_step": int(nbytes_out),
                    "api": api, "steps": steps, "pipeline_depth": depth}
        return e2e_slabs(self, niters, steps, barrier, dist)

    def close(self):
        self.torch.cuda.synchronize()
        capi = self.pkg.capi
        for nb, p in self.peer.items():
            for ptr in p["bufs"]:
                capi.ipc_close(ptr)
            capi.ipc_close(p["flags"])
        self.peer = {}
        if self.dist is not None:
            self.dist.barrier()
        self.t = []
        for m in self.mem:
            m.free()
        if self.scratch is not None:
            self.scratch_t = None
            self.scratch.free()
        if hasattr(self, "flags"):
            self.flags.free()


def e2e_slabs(eng, niters, steps, barrier, dist):
    """End to end at N > 1, one process per GPU: per step every rank uploads ITS slab of every input array from pinned
    host memory (output buffers: boundary shell only, as the C drivers and the N = 1 leg do), runs `niters` sweeps with the
    fused halo push, and downloads its owned part of the result array.  Returns the dict bench.py reports."""
    import os
    import time
    torch = eng.torch
    pkg, info = eng.pkg, eng.info
    # B200_E2E_PIPELINE=2: two engines per rank (two sets of device buffers, peer mappings and flags, each on its own
    # stream) driven alternately, so that one step's copies overlap the other lane's sweeps (N = 2: 31.3 -> 36.4 GLUP/s).
    # Default at N > 1 is ONE lane -- the caller's engine, phase by phase: the two-lane form has only been validated on 2
    # GPUs (its first version dead-locked at N = 8, see step()).
    depth = 2 if os.environ.get("B200_E2E_PIPELINE", "1") == "2" else 1
    lanes = []
    for k in range(depth):
        if depth == 1:
            e, st = eng, eng.stream
        else:
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                e = SlabEngine(pkg, eng.test, eng.real, eng.nx, eng.ny, eng.ns, eng.scalars,
                               world=eng.world, rank=eng.rank, dist=dist, halo=eng.halo, seed=77 + k)
        host = [torch.empty(e._slot_len(q), dtype=e.t[q].dtype).pin_memory() for q in range(info["narrays"])]
        for h in host:
            h.uniform_(-1, 1)
        lanes.append((e, host, st))
    L = eng.layout
    h2d = d2h = 0
    prev_sweeps = None

    def step(i):
        nonlocal h2d, d2h, prev_sweeps
        e, host, st = lanes[i % depth]
        st.synchronize()                      # this lane's previous step has left its host buffers
        e.rewind()
        h2d = e.load_host(host)
        with torch.cuda.stream(st):
            # The sweeps of successive steps run in STEP ORDER on every rank (an event chain between the two lanes): the
            # sweep kernels are persistent and spin on their neighbours' flags, so two ranks executing different lanes'
            # sweeps at the same time would wait for each other for ever (seen at N = 8: a lane's upload finishing early on
            # one rank let its sweeps overtake the other lane's).  Uploads and downloads still overlap the other lane's sweeps.
            if prev_sweeps is not None:
                st.wait_event(prev_sweeps)
            e.run(niters)
            prev_sweeps = torch.cuda.Event()
            prev_sweeps.record(st)
            slot = e.result_slot()
            a, b = (L.own_lo - L.mem_lo) * e.unit, (L.own_hi - L.mem_lo) * e.unit
            host[slot][a:b].copy_(e.t[slot][a:b], non_blocking=True)
            d2h = (b - a) * host[slot].element_size()
        return slot

    def drain():
        for _, _, st in lanes:
            st.synchronize()

    for i in range(depth):
        step(i)
    drain()
    barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    drain()
    barrier()
    dt = (time.perf_counter() - t0) / steps
    if dist is not None:
        tt = torch.tensor([dt, float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dt, h2d_all, d2h_all = float(mx[0].item()), int(tt[1].item()), int(tt[2].item())
    else:
        h2d_all, d2h_all = h2d, d2h
    if depth > 1:
        for e, host, st in lanes:
            e.close()
    return {"seconds_per_step": dt, "h2d_bytes_per_step": h2d_all, "d2h_bytes_per_step": d2h_all,
            "api": "per-rank pinned slab copies: b200_load_shell_slab for output buffers, whole slab otherwise; b200_slab_loop "
                   "(C ABI)" + ("; two engines per rank alternating (copy/compute overlap across steps)" if depth > 1 else ""),
            "steps": steps, "pipeline_depth": depth}


def multi_eq_single(pkg, dist, test, real, nx, ny, ns, scalars, niters, world, rank, halo="push"):
    """Parity carried by the number itself: the N-rank slab run and a 1-GPU run of the SAME global grid (per-GPU extents
    nx x ny x ns, weak-scaled like the bench) on rank 0, from identical deterministic contents, `niters` sweeps each; the
    owned planes of every rank are gathered to rank 0 and compared BIT FOR BIT with the single-GPU arrays.  Returns (on
    rank 0) {"multi_eq_single": bool, "bytes_compared": n, ...}; None on the other ranks.  Outside any timed region."""
    import torch
    info = pkg.test_info(test)
    eng = SlabEngine(pkg, test, real, nx, ny, ns, scalars, world=world, rank=rank, dist=dist, halo=halo)
    eng.fill_pattern()
    dist.barrier()
    eng.run(niters)
    torch.cuda.synchronize()
    L = eng.layout
    slots = sorted({eng.idxs[q] for q in range(info["rotation"])}) if info["rotation"] else list(range(info["narrays"]))
    ok, nbytes, first_bad = True, 0, None
    single = None
    if rank == 0:
        gdims = (nx, ny, ns * world) if info["ndims"] == 3 else (nx, ny * world, 1)
        single = SlabEngine(pkg, test, real, gdims[0], gdims[1], gdims[2], scalars)
        single.fill_pattern()
        single.run(niters)
        torch.cuda.synchronize()
        assert single.idxs == eng.idxs
    for q in slots:
        a, b = (L.own_lo - L.mem_lo) * eng.unit, (L.own_hi - L.mem_lo) * eng.unit
        mine = eng.t[q][a:b].contiguous()
        if rank == 0:
            g = torch.empty(L.n * eng.unit, dtype=mine.dtype, device=mine.device)
            g[: mine.numel()] = mine
            reqs = []
            for r in range(1, world):
                lo, hi = L.n * r // world * eng.unit, L.n * (r + 1) // world * eng.unit
                reqs.append(dist.irecv(g[lo:hi], src=r))
            for rq in reqs:
                rq.wait()
            torch.cuda.synchronize()
            same = torch.equal(g.view(torch.int32 if real == "float" else torch.int64),
                               single.t[q].view(torch.int32 if real == "float" else torch.int64))
            nbytes += g.numel() * g.element_size()
            if not same and first_bad is None:
                first_bad = q
            ok = ok and same
            del g
        else:
            dist.send(mine, dst=0)
    dist.barrier()
    eng.close()
    if single is not None:
        single.close()
    if rank != 0:
        return None
    res = {"multi_eq_single": bool(ok), "bytes_compared": int(nbytes), "test": test, "real": real,
           "global_grid": f"{nx}x{ny}x{ns * world}" if info["ndims"] == 3 else f"{nx}x{ny * world}",
           "niters": niters, "ranks": world, "halo": halo, "slots": slots}
    if first_bad is not None:
        res["first_bad_slot"] = first_bad
    return res


def suite_table(pkg, peak_gbs, full, scalars, niters=10, reps=3, cpu_fn=None):
    """GLUP/s and fraction of the HBM roofline for every test, float and double, on cuda:0."""
    import torch
    rows = []
    # C1 = the README size, C2 = BASELINE configs[2], C3 = the ">= 1024^3 double grid" of the north star
    sizes = [("C1", (512, 256, 256))] + ([("C2", (1024, 1024, 512)), ("C3", (1024, 1024, 1024))] if full else [])
    for label, (nx, ny, ns) in sizes:
        for test in pkg.TESTS:
            if test == "matmul":        # tensor-core bound, not HBM: see matmul_table()
                continue
            info = pkg.test_info(test)
            for real in (("double",) if label == "C3" else ("double", "float")):
                dims = (nx, ny, ns) if info["ndims"] == 3 else (nx, ny * ns, 1)
                try:
                    eng = SlabEngine(pkg, test, real, dims[0], dims[1], dims[2], scalars.get(test, []))
                    eng.run(niters)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(reps):
                        eng.run(niters)
                    e1.record()
                    torch.cuda.synchronize()
                    sec = e0.elapsed_time(e1) * 1e-3 / (reps * niters)
                    lups = eng.local_interior_points()
                    bpl = (info["nread"] + info["nwritten"]) * (4 if real == "float" else 8)
                    rows.append({"test": test, "real": real, "size": f"{dims[0]}x{dims[1]}x{dims[2]}", "cfg": label,
                                 "us_per_sweep": round(sec * 1e6, 2), "glups": round(lups / sec / 1e9, 2),
                                 "gbs": round(lups * bpl / sec / 1e9, 1), "frac": round(lups * bpl / sec / 1e9 / peak_gbs, 4),
                                 "regs": pkg.kernel_info(test, real)["regs"],
                                 "passes_per_10_sweeps": 6 if eng.scratch is not None else 10})
                    eng.close()
                    if cpu_fn is not None and label == "C1":
                        # the gcc path on the box's host cores, same test / size / niters (one timed step)
                        rows[-1].update(cpu_fn(test, real, dims[0], dims[1], dims[2], niters))
                except Exception as e:      # keep the headline alive; report the failure
                    rows.append({"test": test, "real": real, "cfg": label, "error": str(e)[:200]})
                    torch.cuda.synchronize()
    return rows


def matmul_table(pkg, n=8192, reps=3):
    """TFLOP/s of the matmul test (C += A*B, n^3, BASELINE configs[4]) through b200_sweep_loop: the hand-written
    tensor-core kernels; beside them cuBLAS on the same operands, called through torch (addmm_ -- a library baseline timed
    in the bench only; the product library has no library-GEMM path)."""
    import torch
    rows = []
    stream = torch.cuda.current_stream().cuda_stream
    for real, dt in (("double", torch.float64), ("float", torch.float32)):
        try:
            A = torch.rand(n * n, device="cuda", dtype=dt) * 2 - 1
            B = torch.rand(n * n, device="cuda", dtype=dt) * 2 - 1
            row = {"test": "matmul", "real": real, "size": f"{n}x{n}x{n}", "regs": pkg.kernel_info("matmul", real)["regs"]}
            if dt == torch.float32:
                torch.backends.cuda.matmul.allow_tf32 = False          # SGEMM, FP32 accuracy: the like-for-like baseline
            for mode in ("tensor", "cublas"):
                Cm = torch.zeros(n * n, device="cuda", dtype=dt)
                ptrs = [A.data_ptr(), B.data_ptr(), Cm.data_ptr()]
                # column-major C(nx,ns) += A(nx,ny) B(ny,ns)  ==  row-major C^T += B^T A^T
                At, Bt, Ct = A.view(n, n), B.view(n, n), Cm.view(n, n)
                if mode == "tensor":
                    run = lambda k: pkg.capi.sweep_loop("matmul", real, n, n, n, [], ptrs, k, stream=stream)   # noqa: E731
                else:
                    def run(k):
                        for _ in range(k):
                            Ct.addmm_(Bt, At)
                run(1)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run(reps)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                key = "tflops" if mode == "tensor" else "cublas_tflops"
                row[key] = round(2 * n ** 3 / ms / 1e9, 2)
                if mode == "tensor":
                    row["ms_per_sweep"] = round(ms, 3)
                del Cm
            row["vs_cublas"] = round(row["tflops"] / row["cublas_tflops"], 3)
            rows.append(row)
        except Exception as e:
            rows.append({"test": "matmul", "real": real, "error": str(e)[:200]})
            torch.cuda.synchronize()
    return rows
