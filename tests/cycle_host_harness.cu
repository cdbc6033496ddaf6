// TEST INFRASTRUCTURE — runs the host/device bodies of the cyclecheck matching (carpedeam_b200/csrc/cdg_cycle.cuh: part_of, cycle_match_record,
// cycle_match_head) on the CPU, so that the path for very long sub-runs (a k-mer repeated more than `cap` times inside one sequence), which no
// GPU test reaches with the shipped cap of 4096, is exercised with a small cap.  Not a product path: compiled by tests/test_merge_host.py with
// nvcc as a host-only executable; k-mer emission, the sort (std::stable_sort) and the band scan are restated here only to drive those bodies.
// usage: cycle_host_harness <file: u32 n, u32 len[n], bases> <k> <maxSeqLen> <cap>   -> one split diagonal per line
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../carpedeam_b200/csrc/cdg_cycle.cuh"

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    const int k = atoi(argv[2]); const u32 maxSeqLen = (u32) atol(argv[3]), cap = (u32) atol(argv[4]);
    u32 n = 0;
    if (fread(&n, 4, 1, f) != 1) return 2;
    std::vector<u32> len(n);
    if (n && fread(len.data(), 4, n, f) != n) return 2;
    std::vector<u64> off(n + 1, 0), recOff(n + 1, 0), diagOff(n + 1, 0);
    for (u32 s = 0; s < n; s++) {
        off[s + 1] = off[s] + len[s];
        recOff[s + 1] = recOff[s] + ((len[s] < maxSeqLen && len[s] >= (u32) k) ? len[s] - k + 1 : 0);
        diagOff[s + 1] = diagOff[s] + 2 * (len[s] / 3) + 1;
    }
    std::vector<unsigned char> data(off[n] + 1);
    if (off[n] && fread(data.data(), 1, off[n], f) != off[n]) return 2;
    fclose(f);
    const u64 R = recOff[n];
    std::vector<u64> keys(R), vals(R);
    for (u32 s = 0; s < n; s++) {
        if (len[s] >= maxSeqLen || len[s] < (u32) k) continue;
        const unsigned char *q = data.data() + off[s];
        for (u32 p = 0; p + (u32) k <= len[s]; p++) {
            u64 idx = 0;
            for (int i = k - 1; i >= 0; i--) idx = idx * 4ull + (u64) cdg_map_letter(q[p + i]);
            keys[recOff[s] + p] = idx; vals[recOff[s] + p] = ((u64) s << 32) | p;
        }
    }
    std::vector<u64> order(R);
    for (u64 i = 0; i < R; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](u64 a, u64 b) { return keys[a] < keys[b]; });
    std::vector<u64> sk(R), sv(R);
    for (u64 i = 0; i < R; i++) { sk[i] = keys[order[i]]; sv[i] = vals[order[i]]; }
    std::vector<u32> hits(diagOff[n] + 1, 0);
    for (u64 r = 0; r < R; r++) cdgcycle::cycle_match_record(sk.data(), sv.data(), r, len.data(), diagOff.data(), hits.data(), cap);
    for (u64 r = 0; r < R; r++) cdgcycle::cycle_match_head(sk.data(), sv.data(), R, r, len.data(), diagOff.data(), hits.data(), cap);
    for (u32 s = 0; s < n; s++) {   // the band scan of cycle_band_kernel, sequentially
        const u32 L = len[s], third = L / 3;
        u32 found = 0;
        const u32 *h = hits.data() + diagOff[s];
        for (u32 d = 0; L < maxSeqLen && d < 2 * third && !found; d++) {
            if (!h[d]) continue;
            const u32 diag = d + third, diaglen = L - diag, gap = (u32) ((double) diaglen * 0.01);
            const u32 lower = d >= gap ? d - gap : 0u, upper = std::min(d + gap, 2 * third);
            u32 band = 0;
            for (u32 i = lower; i <= upper; i++) if (h[i] <= h[d]) band += h[i];
            const float rate = (float) band / (float) ((u64) diaglen - (u64) k + 1ull);
            if ((double) rate > 0.24) found = diag;
        }
        printf("%u\n", found);
    }
    return 0;
}
