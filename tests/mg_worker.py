"""Multi-GPU parity worker (launched by torchrun, one rank per GPU): the sharded iteration must reproduce the single-GPU
iteration exactly — hits, alignment records and corrected bytes — on the reference's config-1 reads and on synthetic reads."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import carpedeam_b200 as cd
from carpedeam_b200.multigpu import ShardedIteration, TorchComm
from support import GOLD, SeqSet, synth_reads


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = TorchComm(torch.device("cuda", local))
    ctx = cd.Context(local)
    ctx.set_damage(cd.load_damage(os.path.join(GOLD, "dhigh")))
    cases = [("golden nucl_reads k=20", SeqSet.from_golden("nucl_reads"), cd.kmer_params()),
             ("golden contigs k=22 ext", SeqSet.from_golden("assembly_contigs_8_noneCycle"), cd.kmer_params(k=22, include_only_extendable=1)),
             ("synthetic 60k", SeqSet(synth_reads(60000, 300000, 21, n_frac=0.001)), cd.kmer_params())]
    ok = True
    # p2p: streamed peer-memory exchange (routing overlaps the extraction), p2p-counted: count matrix first, nccl: NCCL all-to-all
    modes = os.environ.get("CDG_MG_EXCHANGE", "p2p,p2p-counted,nccl").split(",")
    sharded = {m: ShardedIteration(ctx, comm, exchange=m.split("-")[0]) for m in modes}
    for m in modes:
        if m == "p2p-counted":
            sharded[m].stream_chunks = 0
    for name, s, kp in [(n + " [" + m + "]", s_, k_) for (n, s_, k_) in cases for m in modes]:
        ctx.upload(s.blob, s.off, s.len, s.ext)
        out_range = np.zeros(len(s.blob), np.uint8)
        res = sharded[name[name.rindex("[") + 1:-1]].run(kmer=kp, fetch_hits=True, fetch_alns=True, out_range=out_range)
        parts = [None] * world
        dist.all_gather_object(parts, dict(hits=res["hits"], alns=res["alns"], off=res["range_offset"], nbytes=res["range_bytes"],
                                           out=out_range[:res["range_bytes"]].copy(), changed=res["changed"],
                                           stats={k: res[k] for k in ("records_sent", "records_received", "grouped_sent", "grouped_received", "head_appended")}))
        if rank == 0:
            hits, alns, out, changed = ctx.iteration(kmer=kp, fetch_hits=True)
            for f in ("rep", "tgt", "score", "diag"):
                got = np.concatenate([p["hits"][f] for p in parts])
                if not np.array_equal(got, hits[f]):
                    ok = False; print("MISMATCH", name, "hits", f, len(got), len(hits[f]))
            for f in ("tgt", "keep", "score", "bits", "qs", "qe", "ds", "de", "seqid"):
                got = np.concatenate([p["alns"][f] for p in parts])
                if not np.array_equal(got, alns[f]):
                    ok = False; print("MISMATCH", name, "alns", f)
            merged = s.blob.copy()
            for p in parts:
                merged[p["off"]:p["off"] + p["nbytes"]] = p["out"]
            if not np.array_equal(merged, out) or sum(p["changed"] for p in parts) != changed:
                ok = False; print("MISMATCH", name, "corrected bytes")
            print("case %-36s world=%d hits=%d changed=%d %s" % (name, world, len(hits["rep"]), changed, [p["stats"] for p in parts]))
    # reads block-partitioned over the ranks: every rank packs and uploads only its block (upload_packed_blocks), compact outputs
    for mode in modes:
        s = cases[2][1]
        per = (s.n + world - 1) // world
        lo, hi = min(s.n, rank * per), min(s.n, (rank + 1) * per)
        words, nmask, _ = cd.pack_seqs(s.blob[int(s.off[lo]):int(s.off[hi - 1]) + int(s.len[hi - 1]) + 2], s.off[lo:hi] - s.off[lo], s.len[lo:hi])
        pin = lambda a, dt: torch.from_numpy(a.view(dt)).pin_memory()
        sharded[mode].upload_packed_blocks(pin(words, np.int64), pin(nmask, np.int32), pin(s.len[lo:hi].copy(), np.int32))
        res = sharded[mode].run(kmer=cases[2][2], fetch_alns="compact")
        ex = cd.expand_alns(res["alns_compact"], s.len)
        pn = int(res["patches"].n)
        pent = np.ctypeslib.as_array(res["patches"].entry, shape=(max(1, pn),))[:pn].copy()
        parts = [None] * world
        dist.all_gather_object(parts, dict(ex={k: v.copy() for k, v in ex.items()}, patches=pent, first=int(res["alns_compact"].first_query)))
        if rank == 0:
            ctx.upload(s.blob, s.off, s.len, s.ext)
            hits, alns, out, changed = ctx.iteration(kmer=cases[2][2])
            k = alns["keep"].astype(bool)
            for f in ("tgt", "score", "bits", "qs", "qe", "ds", "de", "seqid"):
                got = np.concatenate([p["ex"][f] for p in parts])
                if not np.array_equal(got, alns[f][k]):
                    ok = False; print("MISMATCH blocks+compact [%s] alns %s" % (mode, f))
            patched = s.blob.copy()
            for p in parts:
                e = p["patches"]
                patched[(e >> np.uint64(8)).astype(np.int64)] = (e & np.uint64(0xFF)).astype(np.uint8)
            if not np.array_equal(patched, out):
                ok = False; print("MISMATCH blocks+compact [%s] corrected bytes" % mode)
            print("case %-36s world=%d kept=%d patches=%d first_query=%s" % ("packed blocks + compact [" + mode + "]", world, int(k.sum()),
                                                                             sum(len(p["patches"]) for p in parts), [p["first"] for p in parts]))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    for sh in sharded.values():
        sh.release()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIGPU PARITY", "OK" if ok else "FAILED")
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
