"""The drop-in boundary inside the reference's own workflow: `ancient_assemble` of the UNMODIFIED reference binary runs with
carpedeam_swap.sh as "$MMSEQS", so createdb, every kmermatcher / rescorediagonal / ancient_correction call of the ten iterations
(data/nuclassemble.sh:105-175), every cyclecheck and the k-mer matcher + Hamming rescoring of the final linclust go to the GPU stage
binary while ancient_read_assemble, ancient_contig_merge, clust, align and the DB utilities stay with the reference.  The final out.fasta must have the SHA-256 the all-reference run produced
(tests/golden/MANIFEST.sha256, written by oracle/make_golden.py) and the intermediate stage DBs must equal the golden ones per key."""
import hashlib
import os
import subprocess

import pytest

from support import GOLD, ROOT, load_golden, read_db

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref", "carpedeam_ref")


def test_reference_workflow_with_gpu_stages_reproduces_out_fasta(tmp_path):
    import carpedeam_b200 as cd
    cd.build()
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/carpedeam_ref (the reference compiled by oracle/build_ref.py) is not present")
    flags = open("/proc/cpuinfo").read()
    if "avx2" not in flags:
        pytest.skip("the reference binary was built for x86-64-v3")
    swap = os.path.join(ROOT, "carpedeam_b200", "host", "carpedeam_swap.sh")
    log = str(tmp_path / "calls.log")
    env = dict(os.environ, CARPEDEAM_REF=REF, CARPEDEAM_GPU=cd.STAGE_BIN, CARPEDEAM_SWAP_LOG=log)
    out = str(tmp_path / "out.fasta")
    r = subprocess.run([swap, "ancient_assemble", os.path.join(GOLD, "test_data.fq.gz"), out, str(tmp_path / "tmp"), "--ancient-damage",
                        os.path.join(GOLD, "dhigh"), "--threads", "8", "--remove-tmp-files", "0", "--delete-tmp-inc", "0"],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    calls = [l.split() for l in open(log)]
    gpu_calls = [c[1] for c in calls if c[0] == "gpu"]
    # ten iterations + the final linclust's k-mer matcher and Hamming pre-clustering (--rescore-mode 0 --wrapped-scoring 1), createdb, five cyclechecks
    assert gpu_calls.count("kmermatcher") == 11 and gpu_calls.count("rescorediagonal") == 11 and gpu_calls.count("ancient_correction") == 10
    assert gpu_calls.count("cyclecheck") == 5 and gpu_calls.count("createdb") == 1
    assert ["ref", "ancient_read_assemble"] in calls and ["ref", "ancient_contig_merge"] in calls and ["ref", "clust"] in calls
    want = dict(l.split()[::-1] for l in open(os.path.join(GOLD, "MANIFEST.sha256")))
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == want["out.fasta"]
    # intermediate DBs of the iterations the golden set holds
    top = os.path.realpath(str(tmp_path / "tmp" / "latest"))
    it = os.path.realpath(os.path.join(top, "nuclassembly_tmp", "latest"))
    for name in ("pref_0", "aln_0", "correction_0", "pref_1", "aln_1", "correction_1", "pref_asm_5", "aln_asm_5", "correction_5",
                 "pref_asm_9", "aln_asm_9", "correction_9", "assembly_reads_4", "assembly_contigs_8_noneCycle"):
        gk, gp, ge = load_golden(name)
        mine = read_db(os.path.join(it, name))
        assert sorted(mine) == gk, name
        bad = [k for k, p, e in zip(gk, gp, ge) if mine[k] != (p, e)]
        assert not bad, (name, len(bad))
