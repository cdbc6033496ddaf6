#!/bin/bash
# Drop-in switch for the reference workflow (data/nuclassemble.sh:85-200): run the assembler with this script in place of the
# reference binary and createdb, the three hot-path modules of the read/contig iterations, cyclecheck and the k-mer matcher + Hamming
# rescoring of the final linclust go to the GPU stage binary; everything else (ancient_read_assemble, ancient_contig_merge, clust,
# align, the DB utilities, ...) stays with the unmodified reference.  The reference exports MMSEQS=argv[0] for its workflow scripts (Application.cpp:198), so the
# reference is exec'ed under this script's name and every "$MMSEQS" <module> call of the workflow comes back here.
#   CARPEDEAM_REF = reference binary, CARPEDEAM_GPU = carpedeam_b200/bin/carpedeam_gpu, CARPEDEAM_SWAP_LOG = optional call log
set -e
: "${CARPEDEAM_REF:?set CARPEDEAM_REF to the reference binary}"
: "${CARPEDEAM_GPU:?set CARPEDEAM_GPU to carpedeam_gpu}"
gpu=0
case "$1" in
    # the hot path (kmermatcher: the assembler's -c 0 calls and linclust's coverage-filtered one; rescorediagonal: --rescore-mode 3 of the
    # iterations and --rescore-mode 0 of linclust's pre-clustering) and the steps either side of it
    kmermatcher|ancient_correction|cyclecheck|createdb) gpu=1 ;;
    rescorediagonal)
        prev=""
        for a in "$@"; do
            if [ "$prev" = "--rescore-mode" ] && { [ "$a" = "3" ] || [ "$a" = "0" ]; }; then gpu=1; fi
            prev="$a"
        done ;;
esac
if [ -n "$CARPEDEAM_SWAP_LOG" ]; then echo "$( [ $gpu = 1 ] && echo gpu || echo ref ) $1" >> "$CARPEDEAM_SWAP_LOG"; fi
if [ $gpu = 1 ]; then exec "$CARPEDEAM_GPU" "$@"; fi
exec -a "$0" "$CARPEDEAM_REF" "$@"
