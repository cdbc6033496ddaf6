"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo processes drive the same Comm adapter the GPU path
uses (variable-split all-to-all ordered by source rank, u64 all-gather, ring shift of the head run) and the ownership
arithmetic that must agree with the device code."""
import os
import subprocess
import sys

import numpy as np

from support import ROOT

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from carpedeam_b200.multigpu import TorchComm, first_run_owner, owner_hash_range, segment_offsets, split_offsets, SENTINEL
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
comm = TorchComm(torch.device("cpu"))
# records: value = source rank * 1e6 + running index, routed by (value %% world); per-source order must survive
n = 1000 + 37 * rank
vals = np.arange(n, dtype=np.int64) + rank * 1_000_000
owner = vals %% world
order = np.argsort(owner, kind="stable")
send = torch.from_numpy(vals[order].copy())
send_counts = [int((owner == o).sum()) for o in range(world)]
recv_counts = comm.exchange_counts(send_counts)
recv = torch.empty(sum(recv_counts), dtype=torch.int64)
comm.all_to_all(recv, send, recv_counts, send_counts)
got = recv.numpy()
off = split_offsets(recv_counts)
for src in range(world):
    seg = got[off[src]:off[src + 1]]
    assert np.all(seg // 1_000_000 == src), "segments are ordered by source rank"
    assert np.all(seg %% world == rank), "every record reached its owner"
    assert np.all(np.diff(seg) > 0), "order inside a source segment is preserved"
assert sum(recv_counts) == sum(int(((np.arange(1000 + 37 * r) + r * 1_000_000) %% world == rank).sum()) for r in range(world))
# peer-memory exchange plan: the count matrix every rank gathers gives the same segment layout an all-to-all produces
rows = comm.all_gather_rows(send_counts + [n, rank])
assert [r[-1] for r in rows] == list(range(world)) and rows[rank][:world] == send_counts
dst, totals = segment_offsets([r[:world] for r in rows], rank)
assert totals[rank] == sum(recv_counts)
assert dst[rank] == off[rank], "own segment starts where the all-to-all put it"
for o in range(world):
    assert dst[o] == sum(rows[r][o] for r in range(rank))
blobs = comm.all_gather_bytes(bytes([rank]) * 72)
assert [b[0] for b in blobs] == list(range(world)) and all(len(b) == 72 for b in blobs)
# u64 all-gather incl. values above 2^63, and the first-run decision
mine = SENTINEL if rank == 1 else (1 << 63) | 12345
keys = comm.all_gather_u64(mine)
assert keys[0] == (1 << 63) | 12345 and keys[1] == SENTINEL
assert first_run_owner(keys) == 0 and first_run_owner([SENTINEL, SENTINEL]) is None and first_run_owner([7, 7]) == 0
# ring shift of the head run: rank r sends its first k records to r-1
buf = torch.arange(10, dtype=torch.int64) + 100 * rank
tail = torch.zeros(8, dtype=torch.int64)
k = 3 + rank
n_recv = comm.shift_to_prev(buf, k, tail)
if rank + 1 < world:
    assert n_recv == 3 + rank + 1 and torch.equal(tail[:n_recv], torch.arange(n_recv) + 100 * (rank + 1))
else:
    assert n_recv == 0
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_comm_adapter_world2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29531", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_hash_ownership_partitions_the_16bit_space():
    from carpedeam_b200.multigpu import owner_hash_range
    for world in (1, 2, 3, 4, 8, 16):
        h = np.arange(65536, dtype=np.uint64)
        dev_owner = (h * np.uint64(world)) >> np.uint64(16)          # owner_of() in cdg_kmer.cuh
        covered = np.zeros(65536, bool)
        for r in range(world):
            lo, hi = owner_hash_range(r, world)
            assert np.all(dev_owner[lo:hi + 1] == r)
            assert not covered[lo:hi + 1].any()
            covered[lo:hi + 1] = True
        assert covered.all()


def test_thread_comm_virtual_ranks_cpu():
    """ThreadComm (virtual ranks as threads, used for hash-range passes on one GPU) must behave like the torch.distributed
    adapter: all-to-all segments ordered by source rank with per-source order preserved, interleaved count exchange, row/bytes
    all-gathers, ring shift of the head run."""
    import threading

    import torch

    from carpedeam_b200.multigpu import LocalWorld, ThreadComm, segment_offsets, split_offsets
    world = 3
    lw = LocalWorld(world)
    errors = []

    def worker(rank):
        try:
            comm = ThreadComm(lw, rank, torch.device("cpu"))
            n = 500 + 11 * rank
            vals = np.arange(n, dtype=np.int64) + rank * 1_000_000
            owner = vals % world
            order = np.argsort(owner, kind="stable")
            send = torch.from_numpy(vals[order].copy())
            send_counts = [int((owner == o).sum()) for o in range(world)]
            ids = [rank * 10 + o for o in range(world)]
            both = comm.exchange_counts(send_counts + ids)
            recv_counts, id_recv = both[0::2], both[1::2]
            assert id_recv == [src * 10 + rank for src in range(world)]
            recv = torch.empty(sum(recv_counts), dtype=torch.int64)
            comm.all_to_all(recv, send, recv_counts, send_counts, async_op=True).wait()
            got = recv.numpy(); off = split_offsets(recv_counts)
            for src in range(world):
                seg = got[off[src]:off[src + 1]]
                assert np.all(seg // 1_000_000 == src) and np.all(seg % world == rank) and np.all(np.diff(seg) > 0)
            rows = comm.all_gather_rows(send_counts)
            dst, totals = segment_offsets(rows, rank)
            assert totals[rank] == sum(recv_counts) and dst[rank] == off[rank]
            assert comm.all_gather_u64((1 << 63) + rank) == [(1 << 63) + r for r in range(world)]
            assert comm.all_gather_bytes(bytes([rank]) * 5) == [bytes([r]) * 5 for r in range(world)]
            buf = torch.arange(10, dtype=torch.int64) + 100 * rank
            tail = torch.zeros(8, dtype=torch.int64)
            n_recv = comm.shift_to_prev(buf, 2 + rank, tail)
            if rank + 1 < world:
                assert n_recv == 3 + rank and torch.equal(tail[:n_recv], torch.arange(n_recv) + 100 * (rank + 1))
            else:
                assert n_recv == 0
            comm.barrier()
        except Exception as e:
            errors.append((rank, repr(e)))
            lw.barrier.abort()

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(60)
    assert not errors, errors


def test_peer_memory_probe_falls_back_collectively():
    """if ONE rank cannot map its peers (CUDA IPC unavailable), EVERY rank must switch to the NCCL exchange: the decision is an
    all-gather of the per-rank outcome.  Driven here with two thread ranks and a stub library in which rank 1 fails."""
    import threading

    import torch

    from carpedeam_b200.multigpu import LocalWorld, ShardedIteration, ThreadComm
    lw = LocalWorld(2)
    out, errors = {}, []

    class StubCtx:
        _h = None

        def _ck(self, rc, what):
            if rc != 0:
                raise RuntimeError(what + " failed")

    class StubLib:
        def __init__(self, rank, fail, where="open"):
            self.rank, self.fail, self.where, self.closed, self.opened = rank, fail, where, [], []

        def cdg_mg_recv_grouped(self, h, total, ptr):
            return -1 if self.fail and self.where == "alloc" else 0

        def cdg_mg_ipc_export(self, h, which, buf):
            if self.fail and self.where == "export":
                return -1
            buf.raw = bytes([self.rank + 1]) * len(buf.raw)
            return 0

        def cdg_mg_ipc_open(self, h, which, peer, buf):
            self.opened.append(peer)
            return -1 if self.fail and self.where == "open" else 0

        def cdg_mg_ipc_close(self, h, which, peer):
            self.closed.append((which, peer))
            return 0

    def worker(rank, fail_rank, where):
        try:
            sh = ShardedIteration.__new__(ShardedIteration)
            sh.ctx, sh.comm, sh.exchange = StubCtx(), ThreadComm(lw, rank, torch.device("cpu")), "p2p"
            sh.L = StubLib(rank, fail=(rank == fail_rank), where=where)
            sh._probe_peer_memory()
            out[(fail_rank, where, rank)] = (sh.exchange, sh.L.closed, sh.L.opened)
        except Exception as e:
            errors.append(repr(e)); lw.barrier.abort()

    # the failure may hit before the handle all-gather (allocation, export) or after it (open): in every case both ranks must pass
    # through the same two collectives and agree
    for fail_rank, where in ((1, "open"), (1, "export"), (0, "alloc"), (None, "open")):
        ts = [threading.Thread(target=worker, args=(r, fail_rank, where)) for r in range(2)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(60)
        assert not errors, (fail_rank, where, errors)
    for where, fr in (("open", 1), ("export", 1), ("alloc", 0)):
        assert out[(fr, where, 0)][0] == "nccl" and out[(fr, where, 1)][0] == "nccl", where     # both ranks, although only one failed
    assert out[(1, "open", 0)][1] == [(4, 1)] and out[(1, "open", 1)][1] == [(4, 0)]      # mappings dropped
    assert out[(1, "export", 0)][2] == []       # a zeroed handle (the peer could not export) is never opened
    assert out[(None, "open", 0)][0] == "p2p" and out[(None, "open", 1)][0] == "p2p" and out[(None, "open", 0)][1] == []
