"""Multi-GPU iteration: one process per GPU, k-mer records owned by hash range (kmermatcher's split scheme,
kmermatcher.cpp:742-784), exchanged with NCCL all-to-all over NVLink through torch.distributed.

Two exchange modes.  "p2p" (default): every rank maps its peers' receive buffers (CUDA IPC over NVLink) and the SENDER's
partition kernel writes each record straight into its owner's memory — the partition pass and the all-to-all are one kernel
(cdg_mg_scatter_records / cdg_mg_scatter_grouped); torch.distributed only moves counts, handles and barriers.  "nccl": partition
locally, then all_to_all_single (the baseline the fused path is measured against).

Plumbing only — every kernel lives in libcarpedeam_gpu.so (cdg_mg_* phases).  The exchange plan is written against a tiny
`Comm` interface so that the host logic (split sizes, ordering by source rank, ring shift of the head run, first-run
decision) is testable on CPU with gloo (tests/test_multigpu_host.py).

Per iteration:
  1. cdg_mg_extract      -> all_to_all of (key, value) records + the slice's identity records  [the data all-to-all north_star names]
  2. cdg_mg_sort         -> all_gather of one u64 per rank (smallest key)  [first-run quirk, kmermatcher.cpp:466]
  3. cdg_mg_group        -> all_to_all of 8-byte grouped records to the rank owning the representative
                            (on-device replacement of the reference's host k-way merge of split files, :953-1111)
  4. cdg_mg_sort2        -> head run (<= 4096 records) to rank-1           [rep-boundary bleed, :886-888]
  5. cdg_mg_finish       -> hits / alignments / corrected bytes of the rank's query range
"""
import ctypes as C

import numpy as np

from . import Alns, CorrParams, Hits, RescoreParams, lib, kmer_params

SENTINEL = 0xFFFFFFFFFFFFFFFF


class CdgErrorLocal(RuntimeError):
    pass


class DevArray:
    """zero-copy view of library-owned device memory for torch (`torch.as_tensor(DevArray(...), device='cuda')`)."""

    def __init__(self, ptr, n, typestr="<i8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def owner_hash_range(rank, world):
    """hash16 values owned by `rank` (owner = hash16 * world >> 16)."""
    lo = (rank * 65536 + world - 1) // world
    hi = ((rank + 1) * 65536 + world - 1) // world - 1
    return lo, hi


def split_offsets(counts):
    off = [0]
    for c in counts:
        off.append(off[-1] + int(c))
    return off


class TorchComm:
    """torch.distributed adapter (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, device):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def exchange_counts(self, send_counts):
        """send_counts: k*world values, [a_0..a_{w-1}, b_0..b_{w-1}, ...] (k count vectors); returns what every peer sends to this
        rank, interleaved per peer: [a from 0, b from 0, a from 1, b from 1, ...] (k = 1: the plain received counts)."""
        k = len(send_counts) // self.world
        t = self.torch.tensor([int(send_counts[j * self.world + o]) for o in range(self.world) for j in range(k)], dtype=self.torch.int64,
                              device=self.device)
        r = self.torch.empty_like(t)
        self.dist.all_to_all_single(r, t)
        return [int(x) for x in r.tolist()]

    class _Synced:
        """wait() of a data collective returns only when the data is in place for EVERY stream: Work.wait() alone orders torch's
        current stream, but the library works on its own non-blocking stream (cdg_stream) — so the host waits for the device."""

        def __init__(self, work, comm):
            self.work, self.comm = work, comm

        def wait(self):
            if self.work is not None:
                self.work.wait()
            self.comm._sync()

    def _sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.current_stream(self.device).synchronize()

    def all_to_all(self, recv, send, recv_counts, send_counts, async_op=False):
        w = self.dist.all_to_all_single(recv, send, output_split_sizes=[int(c) for c in recv_counts],
                                        input_split_sizes=[int(c) for c in send_counts], async_op=async_op)
        if async_op:
            return self._Synced(w, self)
        self._sync()
        return None

    def all_gather_rows(self, row):
        """row: list of ints (int64 range) -> list over ranks of such rows"""
        t = self.torch.tensor([int(x) for x in row], dtype=self.torch.int64, device=self.device)
        out = self.torch.empty(self.world * len(row), dtype=self.torch.int64, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        flat = out.tolist()
        return [flat[r * len(row):(r + 1) * len(row)] for r in range(self.world)]

    def all_gather_bytes(self, payload):
        """payload: bytes of equal length on every rank -> list of bytes per rank"""
        t = self.torch.frombuffer(bytearray(payload), dtype=self.torch.uint8).to(self.device)
        out = self.torch.empty(self.world * len(payload), dtype=self.torch.uint8, device=self.device)
        self.dist.all_gather_into_tensor(out, t)
        b = out.cpu().numpy().tobytes()
        return [b[r * len(payload):(r + 1) * len(payload)] for r in range(self.world)]

    def barrier(self):
        self.dist.barrier()

    def all_gather_u64(self, value):
        t = self.torch.tensor([np.uint64(value).astype(np.int64)], dtype=self.torch.int64, device=self.device)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [int(np.int64(o.item()).astype(np.uint64)) for o in out]

    def shift_to_prev(self, send, n_send, recv_area):
        """rank r sends `send[:n_send]` to rank r-1 and receives from rank r+1; returns the received count.  recv_area: a tensor
        with room for the incoming records, or a callable n -> tensor of n elements (the receiver learns n first)."""
        torch, dist = self.torch, self.dist
        cnt = self.all_gather_u64(n_send)
        n_recv = cnt[self.rank + 1] if self.rank + 1 < self.world else 0
        ops = []
        if self.rank > 0 and n_send > 0:
            ops.append(dist.P2POp(dist.isend, send[:n_send], self.rank - 1))
        if n_recv > 0:
            area = recv_area(n_recv) if callable(recv_area) else recv_area
            ops.append(dist.P2POp(dist.irecv, area[:n_recv], self.rank + 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            self._sync()
        return n_recv


class LocalWorld:
    """shared state of `world` virtual ranks that run as threads of ONE process on ONE GPU (see ThreadComm)"""

    def __init__(self, world):
        import threading
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world


class ThreadComm:
    """Comm adapter for virtual ranks: every rank is a host thread with its own library context on the same device, the
    "collectives" are device-to-device copies between the contexts' buffers.  Two uses: (1) hash-range passes on a single GPU —
    the reference's split mode (kmermatcher.cpp:665-675, setupKmerSplits :742-784) with unsplit results: each virtual rank sorts
    and groups one hash range, so no sort sees more than 1/world of the k-mer records; (2) the multi-rank logic runs in the
    single-GPU test-suite.  Use with ShardedIteration(..., exchange="nccl") (the IPC exchange needs separate processes)."""

    def __init__(self, local_world, rank, device):
        import torch
        self.torch, self.dist, self.device = torch, None, device
        self.lw, self.rank, self.world = local_world, rank, local_world.world

    def _gather(self, obj):
        """every rank contributes `obj`; returns the list over ranks (two barriers: publish, then everybody has read)"""
        self.lw.slots[self.rank] = obj
        self.lw.barrier.wait()
        out = list(self.lw.slots)
        self.lw.barrier.wait()
        return out

    def _sync(self):
        if self.device.type == "cuda":
            self.torch.cuda.synchronize()

    def barrier(self):
        self._sync()
        self.lw.barrier.wait()

    def exchange_counts(self, send_counts):
        k = len(send_counts) // self.world
        rows = self._gather([int(c) for c in send_counts])
        return [rows[src][j * self.world + self.rank] for src in range(self.world) for j in range(k)]

    def all_gather_rows(self, row):
        return self._gather([int(x) for x in row])

    def all_gather_bytes(self, payload):
        return self._gather(bytes(payload))

    def all_gather_u64(self, value):
        return self._gather(int(value))

    class _Done:
        def wait(self):
            return None

    def all_to_all(self, recv, send, recv_counts, send_counts, async_op=False):
        sends = self._gather((send, split_offsets(send_counts)))     # peers' send tensors (same device) and their segment offsets
        pos = 0
        for src in range(self.world):
            t, off = sends[src]
            c = int(recv_counts[src])
            if c:
                recv[pos:pos + c].copy_(t[off[self.rank]:off[self.rank] + c])
            pos += c
        self._sync()
        self.lw.barrier.wait()          # nobody reuses a send buffer before every peer has copied from it
        return self._Done()

    def shift_to_prev(self, send, n_send, recv_area):
        items = self._gather((send, int(n_send)))
        n_recv = 0
        if self.rank + 1 < self.world:
            t, n_recv = items[self.rank + 1]
            if n_recv:
                area = recv_area(n_recv) if callable(recv_area) else recv_area
                area[:n_recv].copy_(t[:n_recv])
        self._sync()
        self.lw.barrier.wait()
        return n_recv


def run_virtual_ranks(device_index, world, blob, off, lens, ext, damage, kmer=None, rescore=None, corr=None):
    """Hash-range passes on one GPU: `world` virtual ranks (threads, one context each) run the sharded iteration with unsplit
    results.  Returns the per-rank result dicts of ShardedIteration.run (hits, alns, corrected bytes of the rank's query range
    in res["out"])."""
    import threading

    import torch

    from . import Context
    lw = LocalWorld(world)
    results, errors = [None] * world, []
    dev = torch.device("cuda", device_index)

    def worker(rank):
        ctx = None
        try:
            torch.cuda.set_device(device_index)
            ctx = Context(device_index)
            ctx.set_damage(damage)
            ctx.upload(blob, off, lens, ext)
            out = np.zeros(len(blob), np.uint8)
            res = ShardedIteration(ctx, ThreadComm(lw, rank, dev), exchange="nccl").run(kmer=kmer, rescore=rescore, corr=corr, fetch_hits=True,
                                                                                    fetch_alns=True, out_range=out)
            res["out"] = out[:res["range_bytes"]].copy()
            results[rank] = res
        except Exception as e:      # a failed rank must not leave the others waiting at a barrier
            errors.append((rank, e))
            lw.barrier.abort()
        finally:
            if ctx is not None:
                ctx.close()

    threads = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        real = [e for e in errors if "BrokenBarrier" not in type(e[1]).__name__] or errors
        raise RuntimeError("virtual rank %d failed: %r" % real[0])
    return results


def segment_offsets(matrix, rank):
    """matrix[r][o] = records rank r sends to rank o.  Returns (dst_off, totals): dst_off[o] = first index of `rank`'s segment in
    o's receive area (segments are ordered by source rank, like an all-to-all's output), totals[o] = records o receives."""
    world = len(matrix)
    dst_off = [sum(int(matrix[r][o]) for r in range(rank)) for o in range(world)]
    totals = [sum(int(matrix[r][o]) for r in range(world)) for o in range(world)]
    return dst_off, totals


def first_run_owner(min_keys):
    """rank holding the globally smallest k-mer key (ties: lowest rank); None if no rank has a record."""
    best, who = None, None
    for r, k in enumerate(min_keys):
        if k == SENTINEL:
            continue
        if best is None or k < best:
            best, who = k, r
    return who


class ShardedIteration:
    """Drives one context (= one GPU) through the five phases."""

    HANDLE_BYTES = 72

    def __init__(self, ctx, comm, exchange="p2p"):
        assert exchange in ("p2p", "nccl")
        self.ctx, self.comm, self.exchange = ctx, comm, exchange
        L = lib()
        L.cdg_mg_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                     C.POINTER(C.c_uint64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.cdg_mg_ipc_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.cdg_mg_ipc_open.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.cdg_mg_ipc_capacity.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
        L.cdg_mg_ipc_close.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.cdg_mg_scatter_records.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.cdg_mg_scatter_grouped.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.cdg_mg_recv_records.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                          C.POINTER(C.c_void_p)]
        L.cdg_mg_sort.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]
        L.cdg_mg_group.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
        L.cdg_mg_recv_grouped.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.cdg_mg_sort2.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.cdg_mg_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                    C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.cdg_mg_head_area.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_void_p)]
        L.cdg_mg_plan.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.cdg_mg_extract_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.cdg_mg_recv_fill.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64]
        L.cdg_mg_finish_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64),
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        self.L = L
        self._up = None      # persistent device staging of upload_sharded_packed
        self._blk = None     # ... and of upload_packed_blocks
        # (records, identity records) the mapped receive areas of the streamed exchange hold: ctx._stream_mapped, shared by every
        # driver object of the context
        self.stream_chunks = 8       # pieces of the slice whose routing overlaps the extraction of the next piece (0: counted exchange)
        if exchange == "p2p" and comm.world > 1:
            self._probe_peer_memory()

    def _probe_peer_memory(self):
        """Collective start-up check: can every rank map every peer's receive buffer (CUDA IPC needs all GPUs visible to all
        processes and a shared IPC namespace)?  If any rank cannot, ALL ranks switch to the NCCL exchange — loudly — instead of
        failing in the middle of the first iteration.  Local steps (allocation, handle export, opening the peers' handles) may
        fail on any rank; the two collectives (handle all-gather, outcome all-gather) are entered by every rank regardless, so a
        failing rank can never leave its peers in a different collective."""
        import sys
        ok, why = 1, ""
        mine = bytes(self.HANDLE_BYTES)
        try:
            r2 = C.c_void_p()
            self.ctx._ck(self.L.cdg_mg_recv_grouped(self.ctx._h, 1, C.byref(r2)), "cdg_mg_recv_grouped")      # allocates the buffer
            h = C.create_string_buffer(self.HANDLE_BYTES)
            self.ctx._ck(self.L.cdg_mg_ipc_export(self.ctx._h, 4, h), "cdg_mg_ipc_export")
            mine = h.raw
        except Exception as e:      # noqa: BLE001 - any failure means "no peer memory here"
            ok, why = 0, str(e)
        handles = self.comm.all_gather_bytes(mine)
        if ok:
            try:
                for peer, blob in enumerate(handles):
                    if peer != self.comm.rank and any(blob):
                        hb = C.create_string_buffer(blob, self.HANDLE_BYTES)
                        self.ctx._ck(self.L.cdg_mg_ipc_open(self.ctx._h, 4, peer, hb), "cdg_mg_ipc_open")
            except Exception as e:      # noqa: BLE001
                ok, why = 0, str(e)
        if not ok:
            print("carpedeam_b200.multigpu: rank %d cannot map peer memory (%s)" % (self.comm.rank, why), file=sys.stderr, flush=True)
        flags = self.comm.all_gather_rows([ok])
        if not all(f[0] for f in flags):
            for peer in range(self.comm.world):
                if peer != self.comm.rank:
                    self.L.cdg_mg_ipc_close(self.ctx._h, 4, peer)
            self.exchange = "nccl"
            if self.comm.rank == 0:
                print("carpedeam_b200.multigpu: peer-memory exchange unavailable, using the NCCL all-to-all exchange", file=sys.stderr, flush=True)

    def _t(self, ptr, n):
        torch = self.comm.torch
        if n == 0:
            return torch.empty(0, dtype=torch.int64, device=self.comm.device)
        return torch.as_tensor(DevArray(ptr, n), device=self.comm.device)

    MIN_BYTES = 8 << 20
    HEAD_RUN = 4096

    def _capacities(self, buffers):
        out = []
        for which in buffers:
            c = C.c_uint64()
            self.ctx._ck(self.L.cdg_mg_ipc_capacity(self.ctx._h, which, C.byref(c)), "cdg_mg_ipc_capacity")
            out.append(int(c.value))
        return out

    def _release_if_growing(self, buffers, needs_by_rank, caps_by_rank):
        """A receive buffer may only be reallocated after every peer dropped its mapping.  All ranks see the same numbers, so they
        take the same branch: if any rank is about to grow a buffer of this group, everyone closes the group and meets at a barrier."""
        grow = any(need > cap for needs, caps in zip(needs_by_rank, caps_by_rank) for need, cap in zip(needs, caps))
        if grow:
            for which in buffers:
                for peer in range(self.comm.world):
                    if peer != self.comm.rank:
                        self.ctx._ck(self.L.cdg_mg_ipc_close(self.ctx._h, which, peer), "cdg_mg_ipc_close")
            self.comm.barrier()
        return grow

    def upload_sharded(self, blob, off, lens, ext=None):
        """Every rank holds the same pinned host arrays (the mapped DB).  Instead of `world` full host-to-device copies, rank r
        copies the r-th 1/world of each array and the ranks all-gather the pieces over NVLink; the library packs from device
        memory.  blob: uint8, off: int64, lens: int32, ext: uint8 or None (torch CPU tensors)."""
        torch, dist, comm = self.comm.torch, self.comm.dist, self.comm
        world, rank, dev = comm.world, comm.rank, comm.device
        import os
        import time
        trace = os.environ.get("CDG_MG_TRACE") is not None
        t0 = time.perf_counter()
        full, h2d = [], 0
        for a in (blob, off, lens, ext):
            if a is None:
                full.append(None)
                continue
            n = a.numel()
            per = (n + world - 1) // world
            lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
            shard = torch.zeros(per, dtype=a.dtype, device=dev)
            shard[:hi - lo].copy_(a[lo:hi], non_blocking=True)
            h2d += (hi - lo) * a.element_size()
            out = torch.empty(per * world, dtype=a.dtype, device=dev)
            dist.all_gather_into_tensor(out, shard)
            full.append(out[:n])
        torch.cuda.current_stream().synchronize()
        t1 = time.perf_counter()
        self.ctx.upload_raw(full[0].data_ptr(), full[0].numel(), full[1].data_ptr(), full[2].data_ptr(),
                            full[3].data_ptr() if full[3] is not None else 0, lens.numel())     # synchronous: `full` may go afterwards
        if trace and rank == 0:
            print("upload_sharded: shards+all_gather %.1f ms, pack from device memory %.1f ms" % ((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3), flush=True)
        return h2d      # bytes this rank copied from the host

    def upload_sharded_packed(self, words, nmask, lens, ext=None):
        """cdg_upload_packed for `world` ranks that all hold the same pinned host arrays (torch CPU tensors: words int64, nmask
        int32 or None, lens int32, ext uint8 or None): rank r copies the r-th 1/world of every array from the host, the ranks
        all-gather the pieces over NVLink into persistent device buffers, the library reads device memory.  Per-rank host traffic
        is 1/world of the packed set; the device buffers are allocated once."""
        torch, dist, comm = self.comm.torch, self.comm.dist, self.comm
        world, rank, dev = comm.world, comm.rank, comm.device
        arrays = (words, nmask, lens, ext)
        sig = tuple(None if a is None else (a.numel(), a.dtype) for a in arrays)
        if self._up is None or self._up[0] != sig:
            bufs = []
            for a in arrays:
                if a is None:
                    bufs.append(None)
                    continue
                per = (a.numel() + world - 1) // world
                bufs.append((per, torch.zeros(per, dtype=a.dtype, device=dev), torch.empty(per * world, dtype=a.dtype, device=dev)))
            self._up = (sig, bufs)
        h2d, full = 0, []
        for a, b in zip(arrays, self._up[1]):
            if a is None:
                full.append(None)
                continue
            per, shard, out = b
            n = a.numel()
            lo, hi = min(n, rank * per), min(n, (rank + 1) * per)
            shard[:hi - lo].copy_(a[lo:hi], non_blocking=True)
            h2d += (hi - lo) * a.element_size()
            dist.all_gather_into_tensor(out, shard)
            full.append(out[:n])
        torch.cuda.current_stream().synchronize()
        self.ctx.upload_packed_raw(full[0].data_ptr(), full[1].data_ptr() if full[1] is not None else 0, full[2].data_ptr(),
                                   full[3].data_ptr() if full[3] is not None else 0, lens.numel(), words.numel())
        return h2d

    def upload_packed_blocks(self, words, nmask, lens, ext=None):
        """Reads block-partitioned over the ranks: every rank holds only ITS block of the packed set in pinned host
        memory (torch CPU tensors: words int64, nmask int32 or None, lens int32, ext uint8 or None; blocks in rank order form the
        set, block r = the rank's query range [r*ceil(n/world), ...)).  Each rank copies its block from the host once, the ranks
        all-gather the blocks over NVLink into persistent device buffers and the library takes the device arrays
        (cdg_upload_packed).  Returns the bytes this rank copied from the host."""
        torch, dist, comm = self.comm.torch, self.comm.dist, self.comm
        world, rank, dev = comm.world, comm.rank, comm.device
        arrays = (words, nmask, lens, ext)
        sig = tuple(None if a is None else (a.numel(), a.dtype) for a in arrays)
        if self._blk is None or self._blk[0] != sig:
            rows = comm.all_gather_rows([lens.numel(), words.numel()] + [0 if a is None else 1 for a in arrays])
            if any(r[2:] != rows[0][2:] for r in rows):
                raise ValueError("upload_packed_blocks: the ranks disagree about which optional arrays exist")
            n_r, w_r = [r[0] for r in rows], [r[1] for r in rows]
            n = sum(n_r)
            per = (n + world - 1) // world
            if any(n_r[r] != max(0, min(n, (r + 1) * per) - min(n, r * per)) for r in range(world)):
                raise ValueError("upload_packed_blocks: block sizes %r are not the query ranges of %d sequences over %d ranks" % (n_r, n, world))
            bufs = []
            for a, cnt in zip(arrays, (w_r, w_r, n_r, n_r)):
                if a is None:
                    bufs.append(None)
                    continue
                mx = max(cnt)
                uniform = all(c == mx for c in cnt)
                gathered = torch.empty(mx * world, dtype=a.dtype, device=dev)
                bufs.append((cnt, mx, torch.zeros(mx, dtype=a.dtype, device=dev), gathered,
                             gathered if uniform else torch.empty(sum(cnt), dtype=a.dtype, device=dev)))
            self._blk = (sig, bufs, n, sum(w_r))
        _, bufs, n, n_words = self._blk
        h2d, full = 0, []
        for a, b in zip(arrays, bufs):
            if a is None:
                full.append(None)
                continue
            cnt, mx, shard, gathered, out = b
            shard[:a.numel()].copy_(a, non_blocking=True)
            h2d += a.numel() * a.element_size()
            dist.all_gather_into_tensor(gathered, shard)
            if out is not gathered:      # ragged blocks: close the gaps (device-to-device)
                pos = 0
                for r in range(world):
                    out[pos:pos + cnt[r]].copy_(gathered[r * mx:r * mx + cnt[r]])
                    pos += cnt[r]
            full.append(out)
        torch.cuda.current_stream().synchronize()
        self.ctx.upload_packed_raw(full[0].data_ptr(), full[1].data_ptr() if full[1] is not None else 0, full[2].data_ptr(),
                                   full[3].data_ptr() if full[3] is not None else 0, n, n_words)
        return h2d

    def release(self):
        """drop every peer mapping (collective): call before the contexts are destroyed or the receive buffers are reused elsewhere"""
        if self.exchange != "p2p":
            return
        self.ctx._stream_mapped = None
        for which in range(5):
            for peer in range(self.comm.world):
                if peer != self.comm.rank:
                    self.ctx._ck(self.L.cdg_mg_ipc_close(self.ctx._h, which, peer), "cdg_mg_ipc_close")
        self.comm.barrier()

    def _map_peers(self, buffers):
        """exchange the IPC handles of this rank's receive buffers `buffers` and map every peer's (cached inside the library)"""
        ctx, comm, L = self.ctx, self.comm, self.L
        mine = b""
        for which in buffers:
            h = C.create_string_buffer(self.HANDLE_BYTES)
            ctx._ck(L.cdg_mg_ipc_export(ctx._h, which, h), "cdg_mg_ipc_export")
            mine += h.raw
        for peer, blob in enumerate(comm.all_gather_bytes(mine)):
            if peer == comm.rank:
                continue
            for i, which in enumerate(buffers):
                hb = C.create_string_buffer(blob[i * self.HANDLE_BYTES:(i + 1) * self.HANDLE_BYTES], self.HANDLE_BYTES)
                ctx._ck(L.cdg_mg_ipc_open(ctx._h, which, peer, hb), "cdg_mg_ipc_open")

    def _phase1_counted(self, send_counts, id_send, rk, rv, irk, irv):
        """counted peer-memory exchange: count matrix first, then one partition pass per rank writes its segments into the peers"""
        ctx, comm, L = self.ctx, self.comm, self.L
        world, rank = comm.world, comm.rank
        # every rank learns the whole count matrix; this all-gather is also the barrier that says every peer is done with
        # the receive buffers of the previous iteration
        rows = comm.all_gather_rows(send_counts + id_send + self._capacities((0, 1, 2, 3)))
        dst, totals = segment_offsets([r[:world] for r in rows], rank)
        id_dst, id_totals = segment_offsets([r[world:2 * world] for r in rows], rank)
        total, id_total = totals[rank], id_totals[rank]
        need = lambda t: max((t + 1) * 8, self.MIN_BYTES)
        if self._release_if_growing((0, 1, 2, 3), [[need(totals[r])] * 2 + [need(id_totals[r])] * 2 for r in range(world)], [r[2 * world:] for r in rows]):
            self.ctx._stream_mapped = None
        ctx._ck(L.cdg_mg_recv_records(ctx._h, total, id_total, C.byref(rk), C.byref(rv), C.byref(irk), C.byref(irv)), "cdg_mg_recv_records")
        self._map_peers((0, 1, 2, 3))
        ctx._ck(L.cdg_mg_scatter_records(ctx._h, (C.c_uint64 * world)(*dst), (C.c_uint64 * world)(*id_dst)), "cdg_mg_scatter_records")
        comm.barrier()      # all partitions have landed
        return total, id_total

    def _phase1_streamed(self, kp, rk, rv, irk, irv, lap):
        """streamed peer-memory exchange (cdg_mg_extract_stream).  Returns (records sent, records in the receive area) or None when
        some rank reported a region overflow (every rank takes the same branch: the flags travel with the counts)."""
        ctx, comm, L = self.ctx, self.comm, self.L
        world, rank = comm.world, comm.rank
        cap = C.c_uint64()
        ctx._ck(L.cdg_mg_plan(ctx._h, C.byref(kp), world, rank, C.byref(cap)), "cdg_mg_plan")
        cap = cap.value
        n = ctx.n
        S, id_cap = world * cap, min(n, 2 * ((n + world - 1) // world) + (1 << 20))
        # the sizes follow from the resident set alone, so every rank sees the same numbers and takes the same branch
        if self.ctx._stream_mapped is None or S > self.ctx._stream_mapped[0] or id_cap > self.ctx._stream_mapped[1]:
            for which in (0, 1, 2, 3):
                for peer in range(world):
                    if peer != rank:
                        ctx._ck(L.cdg_mg_ipc_close(ctx._h, which, peer), "cdg_mg_ipc_close")
            comm.barrier()
            ctx._ck(L.cdg_mg_recv_records(ctx._h, S, id_cap, C.byref(rk), C.byref(rv), C.byref(irk), C.byref(irv)), "cdg_mg_recv_records")
            self._map_peers((0, 1, 2, 3))
            self.ctx._stream_mapped = (S, id_cap)
        else:
            ctx._ck(L.cdg_mg_recv_records(ctx._h, S, id_cap, C.byref(rk), C.byref(rv), C.byref(irk), C.byref(irv)), "cdg_mg_recv_records")
        sc = (C.c_uint64 * world)(); isc = (C.c_uint64 * world)(); ovf = C.c_int()
        ctx._ck(L.cdg_mg_extract_stream(ctx._h, C.byref(kp), self.stream_chunks, cap, sc, isc, C.byref(ovf)), "cdg_mg_extract_stream")
        lap("extract+route (streamed)")
        rows = comm.all_gather_rows(list(sc) + list(isc) + [ovf.value])      # also: every rank's routing kernels have finished
        if any(r[2 * world] for r in rows):
            if rank == 0:
                import sys
                print("carpedeam_b200.multigpu: a receive region would overflow (skewed hash ownership), using the counted exchange", file=sys.stderr, flush=True)
            return None
        ctx._ck(L.cdg_mg_recv_fill(ctx._h, (C.c_uint64 * world)(*[rows[r][rank] for r in range(world)]), cap), "cdg_mg_recv_fill")
        id_dst, id_totals = segment_offsets([r[world:2 * world] for r in rows], rank)
        if id_totals[rank] > id_cap:
            raise CdgErrorLocal("identity receive area too small")      # 2 n / world + 2^20 records: hash ownership is uniform
        ctx._ck(L.cdg_mg_recv_records(ctx._h, S, id_totals[rank], C.byref(rk), C.byref(rv), C.byref(irk), C.byref(irv)), "cdg_mg_recv_records")
        ctx._ck(L.cdg_mg_scatter_records(ctx._h, None, (C.c_uint64 * world)(*id_dst)), "cdg_mg_scatter_records")
        comm.barrier()      # identity records have landed
        return sum(sc), S

    def run(self, kmer=None, rescore=None, corr=None, fetch_hits=False, fetch_alns=False, out_range=None):
        ctx, comm, L = self.ctx, self.comm, self.L
        world, rank = comm.world, comm.rank
        kp = kmer or kmer_params()
        rp = rescore or RescoreParams(0.9, 0.001, 0)
        cp = corr or CorrParams(0.9, 0.99)
        torch = comm.torch
        import time
        tm = {}
        t_last = [time.perf_counter()]

        def lap(name):
            torch.cuda.synchronize()
            now = time.perf_counter()
            tm[name] = tm.get(name, 0.0) + (now - t_last[0]) * 1e3
            t_last[0] = now
        # phase 1
        p2p = self.exchange == "p2p"
        sc = (C.c_uint64 * world)(); isc = (C.c_uint64 * world)()
        sk, sv, ik, iv = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        streamed = p2p and self.stream_chunks > 0 and world > 1
        rk, rv, irk, irv = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        send_counts, id_send, nsend, id_nsend = [], [], 0, 0
        if not streamed:
            ctx._ck(L.cdg_mg_extract(ctx._h, C.byref(kp), world, rank, 0 if p2p else 1, sc, C.byref(sk), C.byref(sv), isc, C.byref(ik), C.byref(iv)),
                    "cdg_mg_extract")
            send_counts = list(sc); id_send = list(isc)
            nsend, id_nsend = sum(send_counts), sum(id_send)
            lap("extract+partition" if not p2p else "extract+count")
        if p2p and self.stream_chunks > 0 and world > 1:
            # streamed exchange: fixed regions per source in every owner's receive area, routing overlaps the extraction
            done = self._phase1_streamed(kp, rk, rv, irk, irv, lap)
            if done is None:      # a region would have overflowed somewhere (degenerate hash ownership): counted exchange instead
                ctx._ck(L.cdg_mg_extract(ctx._h, C.byref(kp), world, rank, 0, sc, C.byref(sk), C.byref(sv), isc, C.byref(ik), C.byref(iv)), "cdg_mg_extract")
                send_counts = list(sc); id_send = list(isc)
                nsend, id_nsend = sum(send_counts), sum(id_send)
                total, id_total = self._phase1_counted(send_counts, id_send, rk, rv, irk, irv)
            else:
                nsend, total = done
        elif p2p:
            total, id_total = self._phase1_counted(send_counts, id_send, rk, rv, irk, irv)
        else:
            both = comm.exchange_counts(send_counts + id_send)
            recv_counts, id_recv = both[0::2], both[1::2]
            total, id_total = sum(recv_counts), sum(id_recv)
            ctx._ck(L.cdg_mg_recv_records(ctx._h, total, id_total, C.byref(rk), C.byref(rv), C.byref(irk), C.byref(irv)), "cdg_mg_recv_records")
            # one exchange: k-mer records and the identity records of the slice (keys and values of each, issued together)
            ws = [comm.all_to_all(self._t(rk.value, total), self._t(sk.value, nsend), recv_counts, send_counts, async_op=True),
                  comm.all_to_all(self._t(rv.value, total), self._t(sv.value, nsend), recv_counts, send_counts, async_op=True),
                  comm.all_to_all(self._t(irk.value, id_total), self._t(ik.value, id_nsend), id_recv, id_send, async_op=True),
                  comm.all_to_all(self._t(irv.value, id_total), self._t(iv.value, id_nsend), id_recv, id_send, async_op=True)]
            for w in ws:
                w.wait()
        lap("all_to_all records")
        # phase 2
        mk = C.c_uint64()
        ctx._ck(L.cdg_mg_sort(ctx._h, C.byref(kp), C.byref(mk)), "cdg_mg_sort")
        lap("identity+sort1")
        owner = first_run_owner(comm.all_gather_u64(mk.value))
        # phase 3
        sc2 = (C.c_uint64 * world)()
        s2 = C.c_void_p()
        ctx._ck(L.cdg_mg_group(ctx._h, C.byref(kp), 1 if owner == rank else 0, 0 if p2p else 1, sc2, C.byref(s2)), "cdg_mg_group")
        send2 = list(sc2)
        lap("group+route" if not p2p else "group+count")
        r2 = C.c_void_p()
        if p2p:
            rows2 = comm.all_gather_rows(send2 + self._capacities((4,)))
            dst2, totals2 = segment_offsets([r[:world] for r in rows2], rank)
            total2 = totals2[rank]
            self._release_if_growing((4,), [[max((totals2[r] + self.HEAD_RUN + 1) * 8, self.MIN_BYTES)] for r in range(world)],
                                     [r[world:] for r in rows2])
            ctx._ck(L.cdg_mg_recv_grouped(ctx._h, total2, C.byref(r2)), "cdg_mg_recv_grouped")
            self._map_peers((4,))
            ctx._ck(L.cdg_mg_scatter_grouped(ctx._h, (C.c_uint64 * world)(*dst2)), "cdg_mg_scatter_grouped")
            comm.barrier()
        else:
            recv2 = comm.exchange_counts(send2)
            total2 = sum(recv2)
            ctx._ck(L.cdg_mg_recv_grouped(ctx._h, total2, C.byref(r2)), "cdg_mg_recv_grouped")
            comm.all_to_all(self._t(r2.value, total2), self._t(s2.value, sum(send2)), recv2, send2)
        lap("all_to_all grouped")
        # phase 4
        sorted_ptr, head_n, head_cap = C.c_void_p(), C.c_uint64(), C.c_uint64()
        ctx._ck(L.cdg_mg_sort2(ctx._h, C.byref(sorted_ptr), C.byref(head_n), C.byref(head_cap)), "cdg_mg_sort2")
        lap("sort2")
        def head_area(n_in):      # the library makes room behind its sorted records (any run length)
            dst = C.c_void_p()
            ctx._ck(L.cdg_mg_head_area(ctx._h, n_in, C.byref(dst)), "cdg_mg_head_area")
            return self._t(dst.value, n_in)
        n_app = comm.shift_to_prev(self._t(sorted_ptr.value, total2), int(head_n.value), head_area)
        lap("head shift")
        # phase 5
        hits, alns = Hits(), Alns()
        off, nbytes, changed = C.c_uint64(), C.c_uint64(), C.c_uint64()
        optr = None
        if out_range is not None:
            optr = C.c_void_p(out_range.ctypes.data) if hasattr(out_range, "ctypes") else C.c_void_p(int(out_range))
        compact, patches = None, None
        if fetch_alns == "compact":      # cdg_alns_compact + cdg_patches of the rank's query range (what crosses PCIe in the benchmark)
            from . import AlnsCompact, Patches
            compact, patches = AlnsCompact(), Patches()
            ctx._ck(L.cdg_mg_finish_compact(ctx._h, C.byref(rp), C.byref(cp), n_app, C.byref(compact), C.byref(patches), C.byref(off), C.byref(nbytes),
                                            C.byref(changed)), "cdg_mg_finish_compact")
        else:
            ctx._ck(L.cdg_mg_finish(ctx._h, C.byref(rp), C.byref(cp), n_app, C.byref(hits) if fetch_hits else None, C.byref(alns) if fetch_alns else None,
                                    optr, C.byref(off), C.byref(nbytes), C.byref(changed)), "cdg_mg_finish")
        lap("reduce+rescore+correct")
        res = dict(phase_ms=tm, range_offset=off.value, range_bytes=nbytes.value, changed=changed.value, records_sent=nsend, records_received=total,
                   grouped_sent=sum(send2), grouped_received=total2, head_appended=n_app)
        if fetch_hits:
            res["hits"] = ctx._hits(hits)
        if fetch_alns == "compact":
            res["alns_compact"], res["patches"] = compact, patches
        elif fetch_alns == "raw":
            res["alns_raw"] = alns     # library-owned pinned arrays, no numpy copies (benchmark)
        elif fetch_alns:
            res["alns"] = ctx._alns(alns)
        return res
