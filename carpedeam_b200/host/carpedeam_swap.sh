#!/bin/bash
# Drop-in switch for the reference workflow (data/nuclassemble.sh:105-175): run the assembler with this script in place of the
# reference binary and the three hot-path modules of the read/contig iterations go to the GPU stage binary, everything else
# (createdb, ancient_read_assemble, ancient_contig_merge, cyclecheck, linclust's own kmermatcher/rescorediagonal calls, ...) stays
# with the unmodified reference.  The reference exports MMSEQS=argv[0] for its workflow scripts (Application.cpp:198), so the
# reference is exec'ed under this script's name and every "$MMSEQS" <module> call of the workflow comes back here.
#   CARPEDEAM_REF = reference binary, CARPEDEAM_GPU = carpedeam_b200/bin/carpedeam_gpu, CARPEDEAM_SWAP_LOG = optional call log
set -e
: "${CARPEDEAM_REF:?set CARPEDEAM_REF to the reference binary}"
: "${CARPEDEAM_GPU:?set CARPEDEAM_GPU to carpedeam_gpu}"
gpu=0
case "$1" in
    ancient_correction) gpu=1 ;;
    kmermatcher|rescorediagonal)
        # the assembler's calls: kmermatcher -c 0, rescorediagonal --rescore-mode 3 (linclust passes -c 0.99 / --rescore-mode 0)
        prev=""
        for a in "$@"; do
            if [ "$1" = kmermatcher ] && [ "$prev" = "-c" ] && [ "$a" = "0" ]; then gpu=1; fi
            if [ "$1" = rescorediagonal ] && [ "$prev" = "--rescore-mode" ] && [ "$a" = "3" ]; then gpu=1; fi
            prev="$a"
        done ;;
esac
if [ -n "$CARPEDEAM_SWAP_LOG" ]; then echo "$( [ $gpu = 1 ] && echo gpu || echo ref ) $1" >> "$CARPEDEAM_SWAP_LOG"; fi
if [ $gpu = 1 ]; then exec "$CARPEDEAM_GPU" "$@"; fi
exec -a "$0" "$CARPEDEAM_REF" "$@"
