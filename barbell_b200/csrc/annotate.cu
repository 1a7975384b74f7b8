// Host engine + C ABI of the annotate hot path (see include/barbell_b200.h).
// One Engine = one CUDA stream with its own device buffers; a bb_ctx owns BB_MAX_INFLIGHT engines so that the
// host->device copy of one batch overlaps the kernels of the previous one (bb_submit / bb_collect).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "host/pack.hpp"
#include "kernels.cuh"

namespace bb {

#define BB_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return BB_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

struct DBuf {   // growable device buffer
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

// ---------------- host-side tables (same alphabet policy S7 as the oracle) ----------------
struct Alphabet {
    uint8_t code[256];
    uint8_t rcchar[256];
    Alphabet() {
        std::memset(code, 0, sizeof code);
        const char* L = "ACGTURYSWKMBDHVN";
        const uint8_t V[] = {1, 2, 4, 8, 8, 5, 10, 6, 9, 12, 3, 14, 13, 11, 7, 15};
        for (int i = 0; L[i]; i++) { code[static_cast<uint8_t>(L[i])] = V[i]; code[static_cast<uint8_t>(L[i] | 0x20)] = V[i]; }
        for (int i = 0; i < 256; i++) rcchar[i] = static_cast<uint8_t>(i);
        const char *a = "ACTGRYSWKMBDHVNX", *b = "TGACYRSWMKVHDBNX";   // reference barcodes.rs:398-441
        for (int i = 0; a[i]; i++) {
            rcchar[static_cast<uint8_t>(a[i])] = static_cast<uint8_t>(b[i]);
            rcchar[static_cast<uint8_t>(a[i] | 0x20)] = static_cast<uint8_t>(b[i] | 0x20);
        }
    }
    static uint8_t comp(uint8_t c) { return static_cast<uint8_t>(((c & 1) << 3) | ((c & 2) << 1) | ((c & 4) >> 1) | ((c & 8) >> 3)); }
};
static const Alphabet kAlpha;

static double lodhi_all_match(int l) {   // reference searcher.rs:229-239 with the recurrence of the reference's forward pass
    volatile double a1 = 0.0, a2 = 0.0, s = 0.0;
    for (int p = 0; p < l; p++) { s = s + 0.5 * a2; a2 = 0.5 * (a2 + a1); a1 = 0.5 * (a1 + 1.0); }
    return s;
}

struct GroupTables {          // device copies for all groups of a ctx (shared by its engines)
    std::vector<DevGroup> host;
    DBuf d_groups, d_blob, d_code;
    int n = 0, max_trace_cols = 0, max_nw = 1, max_region = 0, max_bar_len = 0, max_own_rows = 0, max_k = 0;
    void release() { d_groups.release(); d_blob.release(); d_code.release(); host.clear(); n = 0; }
};

struct Engine {
    int device = 0;
    cudaStream_t stream = nullptr;       // owned stream for bb_annotate / bb_submit
    const GroupTables* gt = nullptr;
    Params prm{};
    std::string err;
    uint64_t launches = 0, h2d_bytes = 0;   // kernels launched / bytes copied host -> device by this engine
    // device buffers
    DBuf d_bases, d_offsets, d_nch, d_chunk_base, d_tile_first, d_tile_span, d_windows, d_entries, d_sorted, d_cub, d_flags, d_hitkeys, d_hits, d_hist, d_rows, d_valid, d_rows_out, d_hits6;
    DBuf d_counters;                     // u32[16]: [0] n_entries [1] entries after unique [2] n_hits [3] n_rows [4..5] kept reads (u64) [6] n_windows
                                         //          [7] filter queue overflow [8] hit list overflow [9] entry slot overflow
    DBuf d_slots, d_slot_cnt, d_nh, d_hit_base;   // slot path: per-read entry slots / counts, reported matches per read, their prefix sum
    uint32_t entries_cap = 0, hits_cap = 0;
    int slot_overflows = 0;
    int glue = 1;                        // 1 = per-read slots, no host round trips (default); 0 = global radix sort path (BB_GLUE=sort; also the fallback)
    bool use_filter = true;              // bb_opts.flags bit 0 disables the pre-filter (exact scan everywhere)
    int pack_mode = 0;                   // bb_opts.flags bit 1: nibble-pack the bases on the host before the PCIe copy (1); bit 2: 2 bits per base
                                         // + an exception list (2; falls back to 1 for good when a batch has too many non-ACGT bytes)
    int pack_threads = 1;
    int pol = 0;                         // bb_opts.policy (barcode_rows.cuh kPol*)
    bool k3_mitm = true;                 // barcode stage keeps half of the traceback records resident (BB_K3_MITM=0: all of them)
    uint8_t* h_pack = nullptr; size_t h_pack_cap = 0;   // pinned staging of the packed bases
    DBuf d_packed;
    uint64_t last_windows = 0;
    uint32_t* h_counters = nullptr;      // pinned, 8 x u32
    bb_row* h_rows = nullptr; size_t h_rows_cap = 0;   // pinned
    cudaEvent_t ev[6] = {};
    cudaEvent_t ev_copy[2] = {};
    double pack_frac = 0.6, pack_rate = 60e9, link_rate = 50e9;   // head fraction packed on the host; bytes/s estimates (see run_host)
    float stage_ms[5] = {0, 0, 0, 0, 0};
    uint32_t last_hits = 0;
    uint64_t last_rows = 0, last_reads = 0, last_kept = 0;

    static double now_ms() {
        static const auto t0 = std::chrono::steady_clock::now();
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    void set_error(const char* fmt, ...) {
        char buf[512];
        va_list ap; va_start(ap, fmt); std::vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
    }
    int init(int dev) {
        device = dev;
        BB_CUDA(cudaSetDevice(device));
        BB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        BB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_counters), 64));
        BB_CUDA(d_counters.ensure(64));
        for (auto& e : ev) BB_CUDA(cudaEventCreate(&e));
        for (auto& e : ev_copy) BB_CUDA(cudaEventCreate(&e));
#if BB_K3_LUT_GLOBAL
        uint32_t lut[256];
        for (int q = 0; q < 256; q++) lut[q] = scan_lut_entry(q);
        BB_CUDA(cudaMemcpyToSymbol(g_k3_scan_lut, lut, sizeof lut));
#endif
        return BB_OK;
    }
    void destroy() {
        cudaSetDevice(device);
        for (DBuf* b : {&d_bases, &d_offsets, &d_nch, &d_chunk_base, &d_tile_first, &d_tile_span, &d_windows, &d_entries, &d_sorted, &d_cub, &d_flags, &d_hitkeys, &d_hits, &d_hist, &d_rows,
                        &d_valid, &d_rows_out, &d_hits6, &d_counters, &d_slots, &d_slot_cnt, &d_nh, &d_hit_base})
            b->release();
        if (h_counters) cudaFreeHost(h_counters);
        if (h_rows) cudaFreeHost(h_rows);
        if (h_pack) cudaFreeHost(h_pack);
        d_packed.release();
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (auto& e : ev_copy) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
    int ensure_host_rows(size_t n) {
        if (n <= h_rows_cap) return BB_OK;
        if (h_rows) cudaFreeHost(h_rows);
        h_rows = nullptr; h_rows_cap = 0;
        size_t want = n + n / 2 + 1024;
        BB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_rows), want * sizeof(bb_row)));
        h_rows_cap = want;
        return BB_OK;
    }

    // K3: one warp (= CTA) per flank match; the grid is as many CTAs as fit the chip at once (shared memory bounds them)
    template <int NWT, bool PACKED, bool S2PAT, bool MITM>
    int launch_barcode_v(const BarArgs& B, uint32_t n_hits, cudaStream_t st) {
        const size_t smem = barcode_rows_smem<NWT, PACKED, MITM>(gt->max_bar_len, gt->max_own_rows);
        static thread_local size_t cfg_smem = 0; static thread_local int per_sm = 0;
        if (cfg_smem != smem) {
            BB_CUDA(cudaFuncSetAttribute(k_barcode_rows<NWT, PACKED, S2PAT, MITM>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            BB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_barcode_rows<NWT, PACKED, S2PAT, MITM>, 32, smem));
            if (per_sm < 1) { set_error("barcode kernel does not fit: %zu bytes of shared memory per warp", smem); return BB_ERR_INVALID; }
            cfg_smem = smem;
        }
        const unsigned blocks = std::min<unsigned>(n_hits, 148u * static_cast<unsigned>(per_sm));
        k_barcode_rows<NWT, PACKED, S2PAT, MITM><<<blocks, 32, smem, st>>>(B);
        launches++;
        BB_CUDA(cudaGetLastError());
        return BB_OK;
    }
    template <int NWT, bool PACKED>
    int launch_barcode(const BarArgs& B, uint32_t n_hits, cudaStream_t st) {
        const bool s2 = (pol & kPolS2PatFirst) != 0;
        if (k3_mitm) return s2 ? launch_barcode_v<NWT, PACKED, true, true>(B, n_hits, st) : launch_barcode_v<NWT, PACKED, false, true>(B, n_hits, st);
        return s2 ? launch_barcode_v<NWT, PACKED, true, false>(B, n_hits, st) : launch_barcode_v<NWT, PACKED, false, false>(B, n_hits, st);
    }

    // The whole device pipeline on stream `st`; inputs resident on the device.
    int run(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, uint64_t total, cudaStream_t st, uint64_t* n_rows) {
        if (!gt || gt->n == 0) { set_error("bb_set_groups has not been called"); return BB_ERR_INVALID; }
        if (n_reads > kMaxBatchReads) { set_error("batch of %u reads exceeds the limit of %u", n_reads, kMaxBatchReads); return BB_ERR_INVALID; }
        if ((reinterpret_cast<uintptr_t>(bases) & 15) != 0) { set_error("bases must be 16-byte aligned"); return BB_ERR_INVALID; }
        BB_CUDA(cudaSetDevice(device));
        last_reads = n_reads; last_rows = 0; last_hits = 0; last_kept = 0;
        *n_rows = 0;
        for (float& f : stage_ms) f = 0.f;
        if (n_reads == 0 || total == 0) return BB_OK;
        bool force_exact = false;
        // slot path unless the slots of this many reads would be out of proportion to the batch (very short reads) or a flank match
        // alone fills a good part of a read's slots (2k + 1 sub-threshold positions, times overlapping windows: large automatic k)
        if (glue == 1 && gt->max_k <= 8 && static_cast<uint64_t>(n_reads) * kSlotCap * 8 <= std::max<uint64_t>(total, 64u << 20)) {
            int redo = 0;
            const int rc = run_slots(bases, offsets, n_reads, total, st, n_rows, &redo);
            if (rc != BB_OK || redo == 0) return rc;
            force_exact = (redo & 1) != 0;           // filter queue overflow: exact scan; slot / hit list overflow: sorted path
            last_rows = 0; last_hits = 0; last_kept = 0; *n_rows = 0;
        }
        return run_sorted(bases, offsets, n_reads, total, st, n_rows, force_exact);
    }

    // chunk index: reads -> chunks of kChunk bases (a chunk never straddles two reads); returns the number of scan tiles
    int chunk_index(const uint64_t* offsets, uint32_t n_reads, uint64_t total, cudaStream_t st, unsigned* n_tiles_out) {
        const uint64_t max_chunks = total / kChunk + n_reads;
        const unsigned n_tiles = static_cast<unsigned>((max_chunks + kScanThreads - 1) / kScanThreads);
        BB_CUDA(d_nch.ensure(static_cast<size_t>(n_reads + 1) * 4));
        BB_CUDA(d_chunk_base.ensure(static_cast<size_t>(n_reads + 1) * 4));
        BB_CUDA(d_tile_first.ensure(static_cast<size_t>(n_tiles) * 4));
        BB_CUDA(d_tile_span.ensure(static_cast<size_t>(n_tiles) * 16));
        k_chunk_count<<<(n_reads + 1 + 255) / 256, 256, 0, st>>>(offsets, n_reads, d_nch.as<uint32_t>());
        launches++;
        {
            size_t tmp = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_nch.as<uint32_t>(), d_chunk_base.as<uint32_t>(), static_cast<int>(n_reads + 1), st);
            BB_CUDA(d_cub.ensure(tmp));
            BB_CUDA(cub::DeviceScan::ExclusiveSum(d_cub.p, tmp, d_nch.as<uint32_t>(), d_chunk_base.as<uint32_t>(), static_cast<int>(n_reads + 1), st));
        }
        k_tile_index<<<(n_tiles + 255) / 256, 256, 0, st>>>(d_chunk_base.as<uint32_t>(), offsets, n_reads, n_tiles, d_tile_first.as<uint32_t>(), d_tile_span.as<uint64_t>());
        launches++;
        BB_CUDA(cudaGetLastError());
        *n_tiles_out = n_tiles;
        return BB_OK;
    }

    // K1 for every group into whatever entry store `A` names; filtered = the pre-filter ran for at least one group
    int launch_scans(ScanArgs A, unsigned n_tiles, uint64_t total, bool force_exact, cudaStream_t st, bool* filtered_out) {
        uint32_t* d_cnt = d_counters.as<uint32_t>();
        bool filtered = false;
        for (int g = 0; g < gt->n; g++) {
            const DevGroup& G = gt->host[g];
            A.group = g; A.strand_xor = (pol & kPolS6RcFirst) ? 1 : 0;
            if (G.f_on && use_filter && !force_exact) {
                // pre-filter (both strands in one 32-bit word) + exact verification of the candidate and read-end windows
                uint32_t win_cap = static_cast<uint32_t>(std::min<uint64_t>(total / 48 + (1u << 20), 1u << 30));
                if (const char* wc = std::getenv("BB_WIN_CAP")) win_cap = static_cast<uint32_t>(std::max(1, std::atoi(wc)));   // test knob: force the overflow fall-back
                BB_CUDA(d_windows.ensure(static_cast<size_t>(win_cap) * 8));
                BB_CUDA(cudaMemsetAsync(d_cnt + 6, 0, 4, st));          // [6] window count ([7] overflow flag is sticky per attempt)
                int halo_l = 0, halo_r = 0;
                filter_halos(G, halo_l, halo_r);
                FilterArgs F{A, d_windows.as<uint64_t>(), d_cnt + 6, win_cap, d_cnt + 7, halo_l, halo_r};
                const size_t smem = filter_smem_bytes(halo_l, halo_r);
                BB_CUDA(cudaFuncSetAttribute(k_flank_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                k_flank_filter<<<n_tiles, kScanThreads, smem, st>>>(F, G);
                launches++;
                VerifyArgs V{A, d_windows.as<uint64_t>(), d_cnt + 6, d_cnt + 7, win_cap};
                if (G.nw == 1) k_flank_verify<1><<<148 * 16, 128, 0, st>>>(V, G);
                else k_flank_verify<2><<<148 * 16, 128, 0, st>>>(V, G);
                launches++;
                filtered = true;
                BB_CUDA(cudaGetLastError());
                continue;
            }
            const size_t smem = 128 + 2 * 256 * G.nw * sizeof(uint64_t) + static_cast<size_t>(kScanThreads) * kChunk + 2 * static_cast<size_t>(G.warm) + 48;
            if (G.nw == 1) {
                BB_CUDA(cudaFuncSetAttribute(k_flank_scan<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                k_flank_scan<1><<<n_tiles, kScanThreads, smem, st>>>(A, G);
            } else {
                BB_CUDA(cudaFuncSetAttribute(k_flank_scan<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                k_flank_scan<2><<<n_tiles, kScanThreads, smem, st>>>(A, G);
            }
            launches++;
            BB_CUDA(cudaGetLastError());
        }
        *filtered_out = filtered;
        return BB_OK;
    }

    // K2b .. K4 + ordered row compaction over the hit list d_hitkeys[0, *(d_cnt + 2)); every kernel reads the count on the device,
    // grids and buffers are sized for `hits_bound` hits
    int launch_tail(const uint8_t* bases, const uint64_t* offsets, uint32_t hits_bound, cudaStream_t st) {
        uint32_t* d_cnt = d_counters.as<uint32_t>();
        BB_CUDA(d_hits.ensure(static_cast<size_t>(hits_bound) * sizeof(Hit)));
        {
            TraceArgs T{};
            const unsigned blocks = std::min<unsigned>((hits_bound + 63) / 64, 148 * 4);
            T.n_slots = blocks * 64;
            BB_CUDA(d_hist.ensure(static_cast<size_t>(gt->max_trace_cols + 2) * 2 * gt->max_nw * 8 * T.n_slots));
            T.bases = bases; T.offsets = offsets; T.hit_keys = d_hitkeys.as<uint64_t>(); T.n_hits = d_cnt + 2;
            T.groups = d_groups(); T.hist = d_hist.as<uint64_t>(); T.hits = d_hits.as<Hit>(); T.pol = pol;
            k_trace<<<blocks, 64, 0, st>>>(T);
            launches++;
            BB_CUDA(cudaGetLastError());
        }
        BB_CUDA(cudaEventRecord(ev[3], st));
        BB_CUDA(d_rows.ensure(static_cast<size_t>(hits_bound) * sizeof(bb_row)));
        BB_CUDA(d_valid.ensure(hits_bound));
        BB_CUDA(cudaMemsetAsync(d_valid.p, 0, hits_bound, st));
        {
            BarArgs B{};
            B.bases = bases; B.offsets = offsets; B.hits = d_hits.as<Hit>(); B.n_hits = d_cnt + 2; B.groups = d_groups();
            B.code = gt->d_code.as<uint8_t>(); B.prm = prm; B.rows = d_rows.as<bb_row>(); B.row_valid = d_valid.as<uint8_t>();
            B.sh_rows = gt->max_bar_len; B.pol = pol;
            // one launch per record format that the geometry can need; each takes the flank matches of its region lengths
            B.rn_lo = -1; B.rn_hi = 48;
            int rc = launch_barcode<1, true>(B, hits_bound, st);
            if (rc != BB_OK) return rc;
            if (gt->max_region > 48) {
                B.rn_lo = 48; B.rn_hi = 64;
                rc = launch_barcode<1, false>(B, hits_bound, st);
                if (rc != BB_OK) return rc;
            }
            if (gt->max_region > 64) {   // large automatic k (custom 115-bp tags)
                B.rn_lo = 64; B.rn_hi = 1 << 20;
                rc = launch_barcode<3, false>(B, hits_bound, st);
                if (rc != BB_OK) return rc;
            }
        }
        BB_CUDA(cudaEventRecord(ev[4], st));
        unsigned long long* d_kept = reinterpret_cast<unsigned long long*>(d_cnt + 4);
        k_collapse<<<(hits_bound + 127) / 128, 128, 0, st>>>(d_hits.as<Hit>(), d_cnt + 2, d_rows.as<bb_row>(), d_valid.as<uint8_t>(), d_kept);
        launches++;
        BB_CUDA(cudaGetLastError());
        BB_CUDA(d_rows_out.ensure(static_cast<size_t>(hits_bound) * sizeof(bb_row)));
        {
            size_t tmp = 0;
            cub::DeviceSelect::Flagged(nullptr, tmp, d_rows.as<bb_row>(), d_valid.as<uint8_t>(), d_rows_out.as<bb_row>(), d_cnt + 3, static_cast<int>(hits_bound), st);
            BB_CUDA(d_cub.ensure(tmp));
            BB_CUDA(cub::DeviceSelect::Flagged(d_cub.p, tmp, d_rows.as<bb_row>(), d_valid.as<uint8_t>(), d_rows_out.as<bb_row>(), d_cnt + 3, static_cast<int>(hits_bound), st));
        }
        return BB_OK;
    }

    // Slot path: no host round trip before the end of the batch.  *redo: bit 0 = the filter's window queue overflowed (re-run with the
    // exact scan), bit 1 = a read had more entries than slots or the batch more matches than the hit list holds (re-run sorted).
    int run_slots(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, uint64_t total, cudaStream_t st, uint64_t* n_rows, int* redo) {
        const uint64_t total16 = (total + 15) & ~15ull;
        uint32_t* d_cnt = d_counters.as<uint32_t>();
        BB_CUDA(cudaEventRecord(ev[0], st));
        unsigned n_tiles = 0;
        int rc = chunk_index(offsets, n_reads, total, st, &n_tiles);
        if (rc != BB_OK) return rc;
        BB_CUDA(d_slots.ensure(static_cast<size_t>(n_reads) * kSlotCap * 8));
        BB_CUDA(d_slot_cnt.ensure(static_cast<size_t>(n_reads + 1) * 4));
        BB_CUDA(d_nh.ensure(static_cast<size_t>(n_reads + 1) * 4));
        BB_CUDA(d_hit_base.ensure(static_cast<size_t>(n_reads + 1) * 4));
        BB_CUDA(cudaMemsetAsync(d_cnt, 0, 64, st));
        BB_CUDA(cudaMemsetAsync(d_slot_cnt.p, 0, static_cast<size_t>(n_reads + 1) * 4, st));
        ScanArgs A{};
        A.bases = bases; A.offsets = offsets; A.n_reads = n_reads; A.total16 = total16;
        A.chunk_base = d_chunk_base.as<uint32_t>(); A.tile_first = d_tile_first.as<uint32_t>(); A.tile_span = d_tile_span.as<uint64_t>();
        A.slots = d_slots.as<uint64_t>(); A.slot_cnt = d_slot_cnt.as<uint32_t>(); A.slot_cap = kSlotCap; A.slot_overflow = d_cnt + 9;
        uint32_t slot_cap = kSlotCap;
        if (const char* sc = std::getenv("BB_SLOT_CAP")) slot_cap = static_cast<uint32_t>(std::min(kSlotCap, std::max(1, std::atoi(sc))));   // test knob
        A.slot_cap = slot_cap; A.slot_stride = kSlotCap;
        bool filtered = false;
        rc = launch_scans(A, n_tiles, total, false, st, &filtered);
        if (rc != BB_OK) return rc;
        BB_CUDA(cudaEventRecord(ev[1], st));
        // per-read resolve (sort + unique + local minima inside one warp), prefix sum of the match counts, gather in read order
        k_read_resolve<<<(n_reads + kResolveWarps - 1) / kResolveWarps, kResolveWarps * 32, 0, st>>>(
            d_slots.as<uint64_t>(), d_slot_cnt.as<uint32_t>(), n_reads, offsets, d_groups(), d_nh.as<uint32_t>(), pol, slot_cap);
        launches++;
        BB_CUDA(cudaMemsetAsync(d_nh.as<uint32_t>() + n_reads, 0, 4, st));
        {
            size_t tmp = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_nh.as<uint32_t>(), d_hit_base.as<uint32_t>(), static_cast<int>(n_reads + 1), st);
            BB_CUDA(d_cub.ensure(tmp));
            BB_CUDA(cub::DeviceScan::ExclusiveSum(d_cub.p, tmp, d_nh.as<uint32_t>(), d_hit_base.as<uint32_t>(), static_cast<int>(n_reads + 1), st));
        }
        if (hits_cap < n_reads * 4ull + 4096) hits_cap = static_cast<uint32_t>(std::min<uint64_t>(n_reads * 4ull + 4096, 1u << 28));
        uint32_t cap = hits_cap;
        if (const char* hc = std::getenv("BB_HITS_CAP")) cap = static_cast<uint32_t>(std::max(1, std::atoi(hc)));   // test knob
        BB_CUDA(d_hitkeys.ensure(static_cast<size_t>(cap) * 8));
        k_hits_gather<<<(n_reads + 255) / 256, 256, 0, st>>>(d_slots.as<uint64_t>(), d_nh.as<uint32_t>(), d_hit_base.as<uint32_t>(), n_reads, cap,
                                                             d_hitkeys.as<uint64_t>(), d_cnt + 2, d_cnt + 8);
        launches++;
        BB_CUDA(cudaGetLastError());
        BB_CUDA(cudaEventRecord(ev[2], st));
        rc = launch_tail(bases, offsets, cap, st);
        if (rc != BB_OK) return rc;
        BB_CUDA(cudaMemcpyAsync(h_counters, d_cnt, 64, cudaMemcpyDeviceToHost, st));
        rc = finish_timing(st, 5);
        if (rc != BB_OK) return rc;
        last_windows = h_counters[6];
        *redo = ((filtered && h_counters[7]) ? 1 : 0) | ((h_counters[8] || h_counters[9]) ? 2 : 0);
        if (h_counters[8]) hits_cap = static_cast<uint32_t>(std::min<uint64_t>(static_cast<uint64_t>(hits_cap) * 4, 1u << 28));
        if (h_counters[9] && ++slot_overflows >= 3) glue = 0;     // repeat-rich input keeps overflowing the slots: stay on the sorted path
        if (*redo) return BB_OK;
        last_hits = h_counters[2];
        last_rows = h_counters[3];
        std::memcpy(&last_kept, h_counters + 4, 8);
        *n_rows = last_rows;
        return BB_OK;
    }

    // Sorted path: global entry list, radix sort, unique, k_resolve, select -- with a host round trip for every count.  Kept for batches
    // of very short reads (slots per read would dwarf the batch), BB_GLUE=sort, and as the fall-back when a slot path capacity overflows.
    int run_sorted(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, uint64_t total, cudaStream_t st, uint64_t* n_rows, bool force_exact_in) {
        const uint64_t total16 = (total + 15) & ~15ull;
        uint32_t* d_cnt = d_counters.as<uint32_t>();
        BB_CUDA(cudaEventRecord(ev[0], st));
        unsigned n_tiles = 0;
        int rc = chunk_index(offsets, n_reads, total, st, &n_tiles);
        if (rc != BB_OK) return rc;

        // ---- K1: flank scan, one launch per group ----
        uint32_t n_entries = 0;
        bool force_exact = force_exact_in, any_filtered = false;
        for (int attempt = 0; attempt < 5; attempt++) {
            if (entries_cap == 0) {
                uint64_t want = std::max<uint64_t>(1u << 20, static_cast<uint64_t>(n_reads) * 32);
                entries_cap = static_cast<uint32_t>(std::min<uint64_t>(want, 1u << 28));
            }
            BB_CUDA(d_entries.ensure(static_cast<size_t>(entries_cap) * 8));
            BB_CUDA(cudaMemsetAsync(d_cnt, 0, 64, st));
            ScanArgs A{};
            A.bases = bases; A.offsets = offsets; A.n_reads = n_reads; A.total16 = total16;
            A.chunk_base = d_chunk_base.as<uint32_t>(); A.tile_first = d_tile_first.as<uint32_t>(); A.tile_span = d_tile_span.as<uint64_t>();
            A.entries = d_entries.as<uint64_t>(); A.n_entries = d_cnt; A.cap = entries_cap;
            rc = launch_scans(A, n_tiles, total, force_exact, st, &any_filtered);
            if (rc != BB_OK) return rc;
            BB_CUDA(cudaMemcpyAsync(h_counters + 6, d_cnt + 6, 8, cudaMemcpyDeviceToHost, st));
            BB_CUDA(cudaMemcpyAsync(h_counters, d_cnt, 4, cudaMemcpyDeviceToHost, st));
            BB_CUDA(cudaStreamSynchronize(st));
            n_entries = h_counters[0];
            last_windows = h_counters[6];
            if (any_filtered && h_counters[7]) { force_exact = true; continue; }   // candidate queue overflow: exact scan instead
            if (n_entries <= entries_cap) break;
            if (attempt == 4 || n_entries > (1u << 28)) { set_error("flank scan produced %u sub-threshold positions; batch too dense", n_entries); return BB_ERR_INVALID; }
            entries_cap = static_cast<uint32_t>(std::min<uint64_t>(static_cast<uint64_t>(n_entries) + n_entries / 4, 1u << 28));
        }
        BB_CUDA(cudaEventRecord(ev[1], st));
        if (n_entries == 0) { return finish_timing(st, 1); }

        // ---- sort entries by (read, group, strand, position) and apply the local-minimum rule ----
        BB_CUDA(d_sorted.ensure(static_cast<size_t>(n_entries) * 8));
        {
            size_t tmp = 0;
            int end_bit = kKeyReadShift;
            while (end_bit < 64 && (static_cast<uint64_t>(n_reads) >> (end_bit - kKeyReadShift)) != 0) end_bit++;
            cub::DeviceRadixSort::SortKeys(nullptr, tmp, d_entries.as<uint64_t>(), d_sorted.as<uint64_t>(), static_cast<int>(n_entries), kKeyPosShift, end_bit, st);
            BB_CUDA(d_cub.ensure(tmp));
            BB_CUDA(cub::DeviceRadixSort::SortKeys(d_cub.p, tmp, d_entries.as<uint64_t>(), d_sorted.as<uint64_t>(), static_cast<int>(n_entries), kKeyPosShift, end_bit, st));
        }
        if (any_filtered) {     // overlapping verification windows emit the same (position, cost) more than once
            size_t tmp = 0;
            uint64_t* uniq = d_entries.as<uint64_t>();     // the unsorted copy is no longer needed
            cub::DeviceSelect::Unique(nullptr, tmp, d_sorted.as<uint64_t>(), uniq, d_cnt + 1, static_cast<int>(n_entries), st);
            BB_CUDA(d_cub.ensure(tmp));
            BB_CUDA(cub::DeviceSelect::Unique(d_cub.p, tmp, d_sorted.as<uint64_t>(), uniq, d_cnt + 1, static_cast<int>(n_entries), st));
            BB_CUDA(cudaMemcpyAsync(h_counters + 1, d_cnt + 1, 4, cudaMemcpyDeviceToHost, st));
            BB_CUDA(cudaStreamSynchronize(st));
            n_entries = h_counters[1];
            BB_CUDA(cudaMemcpyAsync(d_sorted.p, uniq, static_cast<size_t>(n_entries) * 8, cudaMemcpyDeviceToDevice, st));
        }
        BB_CUDA(d_flags.ensure(n_entries));
        h_counters[1] = n_entries;
        BB_CUDA(cudaMemcpyAsync(d_cnt + 1, h_counters + 1, 4, cudaMemcpyHostToDevice, st));   // [1] = entries after the sort / unique
        k_resolve<<<(n_entries + 255) / 256, 256, 0, st>>>(d_sorted.as<uint64_t>(), d_cnt + 1, offsets, d_groups(), d_flags.as<uint8_t>(), pol);
        launches++;
        BB_CUDA(cudaGetLastError());
        BB_CUDA(d_hitkeys.ensure(static_cast<size_t>(n_entries) * 8));
        {
            size_t tmp = 0;
            cub::DeviceSelect::Flagged(nullptr, tmp, d_sorted.as<uint64_t>(), d_flags.as<uint8_t>(), d_hitkeys.as<uint64_t>(), d_cnt + 2, static_cast<int>(n_entries), st);
            BB_CUDA(d_cub.ensure(tmp));
            BB_CUDA(cub::DeviceSelect::Flagged(d_cub.p, tmp, d_sorted.as<uint64_t>(), d_flags.as<uint8_t>(), d_hitkeys.as<uint64_t>(), d_cnt + 2, static_cast<int>(n_entries), st));
        }
        BB_CUDA(cudaMemcpyAsync(h_counters + 2, d_cnt + 2, 4, cudaMemcpyDeviceToHost, st));
        BB_CUDA(cudaStreamSynchronize(st));
        const uint32_t n_hits = h_counters[2];
        last_hits = n_hits;
        BB_CUDA(cudaEventRecord(ev[2], st));
        if (n_hits == 0) { return finish_timing(st, 2); }

        // ---- K2b traceback, K3 barcode stage, K4 collapse + ordered compaction ----
        rc = launch_tail(bases, offsets, n_hits, st);
        if (rc != BB_OK) return rc;
        BB_CUDA(cudaMemcpyAsync(h_counters + 3, d_cnt + 3, 12, cudaMemcpyDeviceToHost, st));
        rc = finish_timing(st, 5);
        if (rc != BB_OK) return rc;
        last_rows = h_counters[3];
        std::memcpy(&last_kept, h_counters + 4, 8);
        *n_rows = last_rows;
        return BB_OK;
    }
    const DevGroup* d_groups() const { return gt->d_groups.as<DevGroup>(); }
    int finish_timing(cudaStream_t st, int n_stages) {
        BB_CUDA(cudaEventRecord(ev[5], st));
        BB_CUDA(cudaStreamSynchronize(st));
        for (int s = 0; s < n_stages && s < 5; s++) {
            cudaEvent_t b = ev[s], e = (s + 1 < n_stages) ? ev[s + 1] : ev[5];
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, b, e) == cudaSuccess) stage_ms[s] = ms;
        }
        return BB_OK;
    }

    // host buffers that already hold the 2-bit wire format (bb_submit_packed): copy, expand on the device, run, rows to pinned memory
    int run_host_packed(const uint8_t* crumbs, uint64_t total, const uint64_t* exc, uint64_t n_exc, const uint64_t* offsets, uint32_t n_reads, uint64_t* n_rows) {
        BB_CUDA(cudaSetDevice(device));
        *n_rows = 0;
        if (n_reads == 0 || total == 0) { last_reads = n_reads; last_rows = 0; last_kept = 0; return BB_OK; }
        static const bool trace = std::getenv("BB_TRACE") != nullptr;        // one line of host timestamps per batch on stderr
        const double tr0 = now_ms();
        const size_t pk = (static_cast<size_t>((total + 3) / 4) + 63) & ~size_t(63);
        BB_CUDA(d_bases.ensure(((total + 15) & ~15ull) + 16));
        BB_CUDA(d_offsets.ensure(static_cast<size_t>(n_reads + 1) * 8));
        BB_CUDA(d_packed.ensure(pk + static_cast<size_t>(n_exc) * 8 + 64));
        BB_CUDA(cudaMemcpyAsync(d_packed.p, crumbs, static_cast<size_t>((total + 3) / 4), cudaMemcpyHostToDevice, stream));
        if (n_exc) BB_CUDA(cudaMemcpyAsync(d_packed.as<uint8_t>() + pk, exc, static_cast<size_t>(n_exc) * 8, cudaMemcpyHostToDevice, stream));
        BB_CUDA(cudaMemcpyAsync(d_offsets.p, offsets, static_cast<size_t>(n_reads + 1) * 8, cudaMemcpyHostToDevice, stream));
        h2d_bytes += (total + 3) / 4 + n_exc * 8 + static_cast<uint64_t>(n_reads + 1) * 8;
        k_unpack_crumbs<<<static_cast<unsigned>((total / 16 + 256) / 256), 256, 0, stream>>>(d_packed.as<uint8_t>(), d_bases.as<uint8_t>(), total);
        launches++;
        if (n_exc) {
            k_patch_exceptions<<<static_cast<unsigned>((n_exc + 255) / 256), 256, 0, stream>>>(
                reinterpret_cast<const uint64_t*>(d_packed.as<uint8_t>() + pk), n_exc, d_bases.as<uint8_t>(), total);
            launches++;
        }
        BB_CUDA(cudaGetLastError());
        const double tr1 = now_ms();
        int rc = run(d_bases.as<uint8_t>(), d_offsets.as<uint64_t>(), n_reads, total, stream, n_rows);
        if (rc != BB_OK) return rc;
        const double tr2 = now_ms();
        if (*n_rows) {
            rc = ensure_host_rows(*n_rows);
            if (rc != BB_OK) return rc;
            BB_CUDA(cudaMemcpyAsync(h_rows, d_rows_out.p, *n_rows * sizeof(bb_row), cudaMemcpyDeviceToHost, stream));
            BB_CUDA(cudaStreamSynchronize(stream));
        }
        if (trace)
            std::fprintf(stderr, "[bb trace] eng %p packed batch: start %.2f queued +%.2f kernels-done +%.2f rows-home +%.2f ms  (%u reads, device stages %.2f ms)\n",
                         static_cast<void*>(this), tr0, tr1 - tr0, tr2 - tr0, now_ms() - tr0, n_reads, stage_ms[0] + stage_ms[1] + stage_ms[2] + stage_ms[3] + stage_ms[4]);
        return BB_OK;
    }
    // bb_reserve: an all-'A' batch of the given shape through the whole path -- every buffer whose size follows from the shape of a batch
    // (all of them on the slot path) is allocated and every kernel is loaded before the first real batch arrives
    int warm(uint32_t n_reads, uint64_t total) {
        BB_CUDA(cudaSetDevice(device));
        if (n_reads == 0 || total == 0) return BB_OK;
        const size_t pk = (static_cast<size_t>((total + 3) / 4) + 63) & ~size_t(63);
        BB_CUDA(d_bases.ensure(((total + 15) & ~15ull) + 16));
        BB_CUDA(d_offsets.ensure(static_cast<size_t>(n_reads + 1) * 8));
        BB_CUDA(d_packed.ensure(pk + static_cast<size_t>(total / 128) * 8 + 64));
        std::vector<uint64_t> off(static_cast<size_t>(n_reads) + 1);
        for (uint32_t i = 0; i <= n_reads; i++) off[i] = static_cast<uint64_t>((static_cast<unsigned __int128>(total) * i) / n_reads);
        BB_CUDA(cudaMemcpyAsync(d_offsets.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, stream));
        BB_CUDA(cudaMemsetAsync(d_packed.p, 0, pk + 64, stream));
        k_unpack_crumbs<<<static_cast<unsigned>((total / 16 + 256) / 256), 256, 0, stream>>>(d_packed.as<uint8_t>(), d_bases.as<uint8_t>(), total);
        k_patch_exceptions<<<1, 256, 0, stream>>>(reinterpret_cast<const uint64_t*>(d_packed.as<uint8_t>() + pk), 0, d_bases.as<uint8_t>(), total);
        BB_CUDA(cudaGetLastError());
        BB_CUDA(cudaStreamSynchronize(stream));                  // (off is read by the copy)
        uint64_t n_rows = 0;
        const uint64_t launches0 = launches;
        int rc = run(d_bases.as<uint8_t>(), d_offsets.as<uint64_t>(), n_reads, total, stream, &n_rows);
        if (rc != BB_OK) return rc;
        rc = ensure_host_rows(static_cast<uint64_t>(n_reads) + n_reads / 2 + 1024);
        if (rc != BB_OK) return rc;
        BB_CUDA(cudaStreamSynchronize(stream));
        launches = launches0; last_reads = 0; last_rows = 0; last_kept = 0; last_hits = 0;   // not a batch of the caller's
        return BB_OK;
    }
    // host-buffer form: copy in, run, copy rows to pinned memory
    int run_host(const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, uint64_t* n_rows) {
        BB_CUDA(cudaSetDevice(device));
        *n_rows = 0;
        if (n_reads == 0) { last_reads = 0; last_rows = 0; last_kept = 0; return BB_OK; }
        static const bool trace = std::getenv("BB_TRACE") != nullptr;        // one line of host timestamps per batch on stderr
        double tr[6] = {now_ms(), 0, 0, 0, 0, 0}, tr_pack_ms = 0;
        const uint64_t total = offsets[n_reads] - offsets[0];
        if (offsets[0] != 0) { set_error("offsets[0] must be 0"); return BB_ERR_INVALID; }
        BB_CUDA(d_bases.ensure(((total + 15) & ~15ull) + 16));
        BB_CUDA(d_offsets.ensure(static_cast<size_t>(n_reads + 1) * 8));
        if (pack_mode && total >= (1u << 20)) {
            // Fewer bytes over PCIe for the HEAD of the batch (packed on the host cores, expanded on the device; lossless for this
            // path) while the TAIL goes over the link as it is: the plain copy is queued first, so the DMA engine moves it while
            // the cores pack.  The split balances the two resources: with P = pack rate, B = link rate (both measured on every
            // batch) and r = packed bytes per base, head fraction x solves (1 - x + r x) / B = x / P.
            const uint64_t split = std::min<uint64_t>(total, static_cast<uint64_t>(pack_frac * static_cast<double>(total))) & ~63ull;
            const size_t n_items = static_cast<size_t>(split >> 22) + 1;                       // work items of the packer (4 M bases each)
            const size_t exc_cap = static_cast<size_t>(split / 64) + (n_items + 1) * 256;       // crumb mode: up to ~1.5 % non-ACGT bytes
            if (h_pack_cap < static_cast<size_t>(total / 2 + 64)) {
                if (h_pack) cudaFreeHost(h_pack);
                h_pack = nullptr; h_pack_cap = 0;
                const size_t want = static_cast<size_t>(total / 2 + 64) + static_cast<size_t>(total / 8) + (1u << 16);
                BB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_pack), want));
                h_pack_cap = want;
            }
            // sized for the whole batch in either format once, not for this batch's split: the split moves from batch to batch and a
            // re-allocation (cudaFree) would stall every stream of the device
            BB_CUDA(d_packed.ensure(static_cast<size_t>(total / 2) + (1u << 16)));
            BB_CUDA(cudaEventRecord(ev_copy[0], stream));
            if (total > split) BB_CUDA(cudaMemcpyAsync(d_bases.as<uint8_t>() + split, bases + split, total - split, cudaMemcpyHostToDevice, stream));
            BB_CUDA(cudaEventRecord(ev_copy[1], stream));
            double pack_s = 0.0;
            size_t wire = 0;                                                                   // bytes of the packed copy
            tr[1] = now_ms();
            bool crumbs = pack_mode == 2 && split > 0;
            if (crumbs) {
                const size_t pk = (static_cast<size_t>(split / 4) + 63) & ~size_t(63);
                size_t n_exc = 0; bool over = false;
                if (pk + exc_cap * 8 > h_pack_cap) over = true;
                else pack_s = pack_crumbs(bases, split, h_pack, reinterpret_cast<uint64_t*>(h_pack + pk), exc_cap, &n_exc, &over, kAlpha.code, pack_threads);
                if (over) { crumbs = false; pack_mode = 1; pack_rate *= 0.5; }                  // N-rich input: nibbles from here on
                else {
                    wire = pk + n_exc * 8;
                    BB_CUDA(cudaMemcpyAsync(d_packed.p, h_pack, wire, cudaMemcpyHostToDevice, stream));
                    k_unpack_crumbs<<<static_cast<unsigned>((split / 16 + 255) / 256), 256, 0, stream>>>(d_packed.as<uint8_t>(), d_bases.as<uint8_t>(), split);
                    launches++;
                    if (n_exc) {
                        k_patch_exceptions<<<static_cast<unsigned>((n_exc + 255) / 256), 256, 0, stream>>>(
                            reinterpret_cast<const uint64_t*>(d_packed.as<uint8_t>() + pk), n_exc, d_bases.as<uint8_t>(), split);
                        launches++;
                    }
                    BB_CUDA(cudaGetLastError());
                }
            }
            if (!crumbs && split) {
                pack_s += pack_nibbles(bases, split, h_pack, kAlpha.code, pack_threads);      // the pool's own time
                wire = static_cast<size_t>(split / 2);
                BB_CUDA(cudaMemcpyAsync(d_packed.p, h_pack, wire, cudaMemcpyHostToDevice, stream));
                k_unpack_nibbles<<<static_cast<unsigned>((split / 16 + 255) / 256), 256, 0, stream>>>(d_packed.as<uint8_t>(), d_bases.as<uint8_t>(), split);
                launches++;
                BB_CUDA(cudaGetLastError());
            }
            tr[2] = now_ms(); tr_pack_ms = pack_s * 1e3;
            BB_CUDA(cudaMemcpyAsync(d_offsets.p, offsets, static_cast<size_t>(n_reads + 1) * 8, cudaMemcpyHostToDevice, stream));
            BB_CUDA(cudaEventSynchronize(ev_copy[1]));            // run() synchronises the stream a few kernels later anyway
            tr[3] = now_ms();
            float copy_ms = 0.f;
            // P = rate of the (shared) packing pool while it works; B = the link's rate = the fastest tail copy seen lately (a copy
            // that queued behind another batch's copy looks slower than the link is)
            if (total - split >= (1u << 20) && cudaEventElapsedTime(&copy_ms, ev_copy[0], ev_copy[1]) == cudaSuccess && copy_ms > 0.f)
                link_rate = std::max(0.98 * link_rate, static_cast<double>(total - split) / (copy_ms * 1e-3));
            if (split >= (1u << 20) && pack_s > 0.0) pack_rate = 0.5 * pack_rate + 0.5 * (static_cast<double>(split) / pack_s);
            h2d_bytes += (total - split) + wire + static_cast<uint64_t>(n_reads + 1) * 8;
            const double saved = pack_mode == 2 ? 0.75 : 0.5;     // 1 - r
            const double x = pack_rate / (link_rate + saved * pack_rate);
            pack_frac = std::min(1.0, std::max(0.05, 0.5 * pack_frac + 0.5 * x));
            static const char* fixed = std::getenv("BB_PACK_FRAC");                           // experiment knob: pin the split
            if (fixed) pack_frac = std::atof(fixed);
        } else {
            BB_CUDA(cudaMemcpyAsync(d_bases.p, bases, total, cudaMemcpyHostToDevice, stream));
            BB_CUDA(cudaMemcpyAsync(d_offsets.p, offsets, static_cast<size_t>(n_reads + 1) * 8, cudaMemcpyHostToDevice, stream));
            h2d_bytes += total + static_cast<uint64_t>(n_reads + 1) * 8;
        }
        int rc = run(d_bases.as<uint8_t>(), d_offsets.as<uint64_t>(), n_reads, total, stream, n_rows);
        if (rc != BB_OK) return rc;
        tr[4] = now_ms();
        if (*n_rows) {
            rc = ensure_host_rows(*n_rows);
            if (rc != BB_OK) return rc;
            BB_CUDA(cudaMemcpyAsync(h_rows, d_rows_out.p, *n_rows * sizeof(bb_row), cudaMemcpyDeviceToHost, stream));
            BB_CUDA(cudaStreamSynchronize(stream));
        }
        tr[5] = now_ms();
        if (trace)
            std::fprintf(stderr, "[bb trace] eng %p start %.2f tail-queued +%.2f packed(+wait) +%.2f tail-arrived +%.2f kernels-done +%.2f rows-home +%.2f  frac %.2f pack %.2f ms\n",
                         static_cast<void*>(this), tr[0], tr[1] - tr[0], tr[2] - tr[0], tr[3] - tr[0], tr[4] - tr[0], tr[5] - tr[0], pack_mode ? pack_frac : 0.0, tr_pack_ms);
        return BB_OK;
    }
};

}  // namespace bb

// ---------------------------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------------------------
struct bb_job {
    const uint8_t* bases; const uint64_t* offsets; uint32_t n_reads; uint64_t tag; int engine;
    int rc = 0; uint64_t n_rows = 0; bool done = false;
    bool packed = false; uint64_t n_bases = 0; const uint64_t* exc = nullptr; uint64_t n_exc = 0;   // bb_submit_packed: `bases` holds crumbs
    bool warm = false;                                                                               // bb_reserve: no data, n_reads x n_bases only
};

struct bb_ctx {
    bb_opts opts{};
    bb::GroupTables gt;
    bb::Engine eng[BB_MAX_INFLIGHT];
    std::string err;
    uint64_t total_reads = 0, kept_reads = 0;
    // pipelined mode: one worker thread per engine
    std::thread workers[BB_MAX_INFLIGHT];
    std::mutex mu;
    std::condition_variable cv;
    std::deque<bb_job*> queue[BB_MAX_INFLIGHT];   // per engine FIFO
    std::deque<bb_job*> order;                     // submission order
    bool stop = false, workers_started = false;
    uint64_t submitted = 0;
    int last_engine = 0;
};

static void ctx_error(bb_ctx* c, const std::string& s) { c->err = s; }

static void worker_main(bb_ctx* c, int idx) {
    for (;;) {
        bb_job* job = nullptr;
        {
            std::unique_lock<std::mutex> lk(c->mu);
            c->cv.wait(lk, [&] { return c->stop || !c->queue[idx].empty(); });
            if (c->queue[idx].empty()) return;
            job = c->queue[idx].front();
        }
        uint64_t n_rows = 0;
        const int rc = job->warm ? c->eng[idx].warm(job->n_reads, job->n_bases)
                     : job->packed ? c->eng[idx].run_host_packed(job->bases, job->n_bases, job->exc, job->n_exc, job->offsets, job->n_reads, &n_rows)
                                   : c->eng[idx].run_host(job->bases, job->offsets, job->n_reads, &n_rows);
        {
            std::lock_guard<std::mutex> lk(c->mu);
            job->rc = rc; job->n_rows = n_rows; job->done = true;
            c->queue[idx].pop_front();
        }
        c->cv.notify_all();
    }
}

extern "C" {

int bb_create(const bb_opts* opts, bb_ctx** out, char* err, size_t errlen) {
    auto fail = [&](int code, const std::string& msg) { if (err && errlen) std::snprintf(err, errlen, "%s", msg.c_str()); return code; };
    if (!opts || !out) return fail(BB_ERR_INVALID, "null argument");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return fail(BB_ERR_NO_DEVICE, std::string("no CUDA device available (") + cudaGetErrorString(e) + "); barbell_b200 has no CPU path");
    if (opts->device < 0 || opts->device >= n_dev) return fail(BB_ERR_INVALID, "device ordinal out of range");
    if (!(opts->alpha > 0.0f) || !(opts->alpha <= 1.0f)) return fail(BB_ERR_INVALID, "alpha must be in (0, 1]");
    cudaDeviceProp prop{};
    cudaGetDeviceProperties(&prop, opts->device);
    if (prop.major < 10) return fail(BB_ERR_NO_DEVICE, "barbell_b200 is built for sm_100a (B200) only");
    auto* c = new bb_ctx();
    c->opts = *opts;
    for (int i = 0; i < BB_MAX_INFLIGHT; i++) {
        int rc = c->eng[i].init(opts->device);
        if (rc != BB_OK) { std::string m = c->eng[i].err; for (int j = 0; j <= i; j++) c->eng[j].destroy(); delete c; return fail(rc, m); }
        c->eng[i].gt = &c->gt;
        c->eng[i].prm.min_score = opts->min_score; c->eng[i].prm.min_score_diff = opts->min_score_diff;
        c->eng[i].use_filter = (opts->flags & 1u) == 0;
        c->eng[i].pack_mode = (opts->flags & 4u) ? 2 : (opts->flags & 2u) ? 1 : 0;
        c->eng[i].pack_threads = bb::pack_default_threads();
        c->eng[i].pol = static_cast<int>(opts->policy) & bb::kPolMask;
        if (const char* e = std::getenv("BB_K3_MITM")) c->eng[i].k3_mitm = std::atoi(e) != 0;
        if (const char* e = std::getenv("BB_GLUE")) c->eng[i].glue = std::strcmp(e, "sort") == 0 ? 0 : 1;
    }
    *out = c;
    return BB_OK;
}

void bb_destroy(bb_ctx* c) {
    if (!c) return;
    if (c->workers_started) {
        { std::lock_guard<std::mutex> lk(c->mu); c->stop = true; }
        c->cv.notify_all();
        for (auto& w : c->workers) if (w.joinable()) w.join();
        for (auto* j : c->order) delete j;
    }
    cudaSetDevice(c->opts.device);
    for (auto& e : c->eng) e.destroy();
    c->gt.release();
    delete c;
}

const char* bb_last_error(const bb_ctx* c) { return c ? c->err.c_str() : "null ctx"; }

int bb_set_groups(bb_ctx* c, const bb_group* groups, int32_t n_groups) {
    using namespace bb;
    if (!c || !groups || n_groups <= 0) return BB_ERR_INVALID;
    auto bad = [&](const std::string& m) { ctx_error(c, m); return BB_ERR_INVALID; };
    if (n_groups > kMaxGroups) return bad("at most 8 query groups are supported");
    if (cudaSetDevice(c->opts.device) != cudaSuccess) return bad("cudaSetDevice failed");
    const float alpha = c->opts.alpha;
    // blob layout per group: eq[2][256][nw] u64 | eq_top | filter masks | bar_off[2][nb][64] u8 | sh_off[2][64] u8 | ov[m+1] i32 (padded to 8)
    const int pol = static_cast<int>(c->opts.policy) & kPolMask;
    std::vector<uint64_t> blob;
    std::vector<size_t> off_eq(n_groups), off_eqt(n_groups), off_feq(n_groups), off_bar(n_groups), off_ov(n_groups);
    std::vector<DevGroup> hg(n_groups);
    int max_trace = 0, max_nw = 1, max_region = 0, max_bar_len = 0, max_own_rows = 0;
    auto over_cost = [&](int t) {             // policy S3: floor (default) / round-to-nearest / ceil of the f32 product
        const float v = static_cast<float>(t) * alpha;
        return (pol & kPolS3Round) ? static_cast<int>(std::floor(v + 0.5f)) : (pol & kPolS3Ceil) ? static_cast<int>(std::ceil(v)) : static_cast<int>(std::floor(v));
    };
    for (int g = 0; g < n_groups; g++) {
        const bb_group& S = groups[g];
        DevGroup& D = hg[g];
        std::memset(&D, 0, sizeof D);
        const int m = S.flank_len;
        if (m < 1 || m > 64 * kMaxFlankWords) return bad("flank length must be 1..128");
        if (S.bar_len < 1 || S.bar_len > kMaxBarLen) return bad("padded barcode length must be 1..64");
        if (S.n_barcodes < 1 || S.n_barcodes > 32 * kMaxBarRounds) return bad("1..4096 barcodes per group");
        if (S.k_flank < 0 || S.k_flank > 120) return bad("flank threshold must be 0..120");
        const int ov_m = over_cost(m);
        if (S.bar1 - S.bar0 + 1 + S.k_flank + 2 * kPadding > kRegionMax) return bad("barcode region (mask + k + 20) exceeds 160 characters");
        if (S.bar0 < 0 || S.bar1 < S.bar0 || S.bar1 >= m || S.pad0 < 0 || S.pad0 > S.bar0) return bad("inconsistent bar/pad regions");
        D.m = m; D.nw = (m + 63) / 64; D.last_bit = (m - 1) & 63; D.k = S.k_flank;
        D.bar0 = S.bar0; D.bar1 = S.bar1; D.pad0 = S.pad0; D.pad1 = S.pad1;
        D.bar_len = S.bar_len; D.n_barcodes = S.n_barcodes; D.match_type = S.match_type;
        D.k_bar = static_cast<int>(static_cast<float>(S.bar_len) * 0.4f);
        D.pbar0 = S.bar0 - S.pad0; D.pbar1 = S.bar1 - S.pad0;
        D.ov_m = ov_m;
        D.warm = ((m + S.k_flank + kGroup - 1) / kGroup) * kGroup;
        D.halo = (D.warm + 15) & ~15;
        D.trace_cols = m + 2 * std::min(S.k_flank, m) + 8;
        D.perfect = lodhi_all_match(S.pad1 - S.pad0);
        max_trace = std::max(max_trace, D.trace_cols); max_nw = std::max(max_nw, D.nw);
        max_region = std::max(max_region, S.bar1 - S.bar0 + 1 + S.k_flank + 2 * kPadding);
        max_bar_len = std::max(max_bar_len, S.bar_len);
        std::vector<uint8_t> pc(m);
        for (int i = 0; i < m; i++) pc[i] = kAlpha.code[static_cast<uint8_t>(S.flank[i])];
        std::vector<int> ov(m + 1);
        for (int t = 0; t <= m; t++) ov[t] = over_cost(t);
        for (int i = 0; i < m; i++) {
            D.pv_plain[i >> 6] |= 1ull << (i & 63);
            if (ov[i + 1] - ov[i]) D.pv_over[i >> 6] |= 1ull << (i & 63);
        }
        off_eq[g] = blob.size();
        blob.resize(blob.size() + 2 * 256 * D.nw, 0);
        for (int s = 0; s < 2; s++)
            for (int ch = 0; ch < 256; ch++) {
                uint8_t code = kAlpha.code[ch];
                if (s == 1) code = Alphabet::comp(code);
                for (int i = 0; i < m; i++)
                    if (pc[i] & code) blob[off_eq[g] + (static_cast<size_t>(s) * 256 + ch) * D.nw + (i >> 6)] |= 1ull << (i & 63);
            }
        // top-aligned copy for the scan kernel: row i at bit i + shift, wildcard rows (all ones) below
        off_eqt[g] = blob.size();
        blob.resize(blob.size() + 2 * 256 * D.nw, 0);
        {
            const int shift = 64 * D.nw - m;
            for (int i = 0; i < m; i++) {
                const int bpos = i + shift;
                D.pv_plain_top[bpos >> 6] |= 1ull << (bpos & 63);
                if (ov[i + 1] - ov[i]) D.pv_over_top[bpos >> 6] |= 1ull << (bpos & 63);
            }
            for (int s = 0; s < 2; s++)
                for (int ch = 0; ch < 256; ch++) {
                    uint64_t* dst = &blob[off_eqt[g] + (static_cast<size_t>(s) * 256 + ch) * D.nw];
                    const uint64_t* src = &blob[off_eq[g] + (static_cast<size_t>(s) * 256 + ch) * D.nw];
                    for (int i = 0; i < m; i++)
                        if ((src[i >> 6] >> (i & 63)) & 1ull) dst[(i + shift) >> 6] |= 1ull << ((i + shift) & 63);
                    for (int bpos = 0; bpos < shift; bpos++) dst[bpos >> 6] |= 1ull << (bpos & 63);
                }
        }
        // pre-filter: the longest N-free run of the flank, at most 15 rows; enabled when 3k <= rows (selective enough)
        off_feq[g] = blob.size();
        blob.resize(blob.size() + 256, 0);
        {
            int best0 = 0, bestn = 0, cur0 = 0, curn = 0;
            for (int i = 0; i <= m; i++) {
                if (i < m && pc[i] != 15 && pc[i] != 0) { if (!curn) cur0 = i; curn++; }
                else { if (curn > bestn) { bestn = curn; best0 = cur0; } curn = 0; }
            }
            const int q = std::min(bestn, 15);
            D.f_q = q; D.f_q0 = best0;
            D.f_on = (q >= 8 && 3 * S.k_flank <= q) ? 1 : 0;
            // second N-free run (outside the first) for the pre-check of the candidates: at most 15 rows, taken next to the mask
            int s_best0 = 0, s_bestn = 0; cur0 = 0; curn = 0;
            for (int i = 0; i <= m; i++) {
                const bool inside_q = i >= best0 && i < best0 + q;
                if (i < m && !inside_q && pc[i] != 15 && pc[i] != 0) { if (!curn) cur0 = i; curn++; }
                else { if (curn > s_bestn) { s_bestn = curn; s_best0 = cur0; } curn = 0; }
            }
            D.f_qs = s_bestn >= 4 ? std::min(s_bestn, 15) : 0;
            D.f_s0 = (s_best0 > best0) ? s_best0 : s_best0 + (s_bestn - std::min(s_bestn, 15));   // keep the rows nearest to the first run
            if (D.f_qs) {
                uint32_t* seq = reinterpret_cast<uint32_t*>(blob.data() + off_feq[g]) + 256;
                for (int ch = 0; ch < 256; ch++) {
                    const uint8_t code = kAlpha.code[ch], ccode = Alphabet::comp(code);
                    uint32_t v = 0;
                    for (int r = 0; r < D.f_qs; r++) {
                        if (pc[D.f_s0 + r] & code) v |= 1u << r;
                        if (pc[D.f_s0 + D.f_qs - 1 - r] & ccode) v |= 1u << (16 + r);
                    }
                    seq[ch] = v;
                }
            }
            uint32_t* feq = reinterpret_cast<uint32_t*>(blob.data() + off_feq[g]);
            for (int ch = 0; ch < 256 && q > 0; ch++) {
                const uint8_t code = kAlpha.code[ch], ccode = Alphabet::comp(code);
                uint32_t v = 0;
                for (int r = 0; r < q; r++) {
                    if (pc[best0 + r] & code) v |= 1u << r;                       // run, forward strand
                    if (pc[best0 + q - 1 - r] & ccode) v |= 1u << (16 + r);       // reverse complement of the run
                }
                feq[ch] = v;
            }
        }
        // barcode patterns as one byte per row = 8 * (4-bit IUPAC set): forward, and explicitly reverse-complemented (barcodes.rs:85-90);
        // + per strand the leading rows all barcodes share (computed once per flank match by k_barcode_rows)
        if (blob.size() & 1) blob.push_back(0);          // the kernel copies the codes with 16-byte loads
        off_bar[g] = blob.size();
        const int n_rounds = (S.n_barcodes + 31) / 32;
        blob.resize(blob.size() + (static_cast<size_t>(2) * n_rounds * 64 * 32 + 2 * 64) / 8, 0);
        {
            // table layout [strand][round][row][lane]: a warp's load of one pattern row is 32 consecutive bytes
            uint8_t* off = reinterpret_cast<uint8_t*>(blob.data() + off_bar[g]);
            uint8_t* shoff = off + static_cast<size_t>(2) * n_rounds * 64 * 32;
            auto at = [&](int st, int b, int row) -> uint8_t& { return off[((static_cast<size_t>(st) * n_rounds + (b >> 5)) * 64 + row) * 32 + (b & 31)]; };
            for (int b = 0; b < S.n_barcodes; b++)
                for (int i = 0; i < S.bar_len; i++) {
                    const uint8_t ch = static_cast<uint8_t>(S.barcodes[static_cast<size_t>(b) * S.bar_len + i]);
                    at(0, b, i) = static_cast<uint8_t>(kAlpha.code[ch] << 3);
                    at(1, b, S.bar_len - 1 - i) = static_cast<uint8_t>(kAlpha.code[kAlpha.rcchar[ch]] << 3);
                }
            for (int st = 0; st < 2; st++) {
                int P = S.bar_len;
                for (int b = 1; b < S.n_barcodes; b++) {
                    int q = 0;
                    while (q < P && at(st, b, q) == at(st, 0, q)) q++;
                    P = q;
                }
                D.sh_p[st] = P;
                for (int q = 0; q < 64; q++) shoff[64 * st + q] = at(st, 0, q);
                max_own_rows = std::max(max_own_rows, S.bar_len - P);
            }
        }
        off_ov[g] = blob.size();
        blob.resize(blob.size() + (m + 2) / 2 + 1, 0);
        std::memcpy(reinterpret_cast<int*>(blob.data() + off_ov[g]), ov.data(), sizeof(int) * (m + 1));
    }
    bb::GroupTables& T = c->gt;
    for (auto& e : c->eng) cudaStreamSynchronize(e.stream);
    if (T.d_blob.ensure(blob.size() * 8) != cudaSuccess || T.d_groups.ensure(sizeof(DevGroup) * n_groups) != cudaSuccess ||
        T.d_code.ensure(256) != cudaSuccess)
        return (ctx_error(c, "cudaMalloc failed for the pattern tables"), BB_ERR_CUDA);
    const uint64_t* base = T.d_blob.as<uint64_t>();
    for (int g = 0; g < n_groups; g++) {
        hg[g].eq = base + off_eq[g];
        hg[g].eq_top = base + off_eqt[g];
        hg[g].f_eq = reinterpret_cast<const uint32_t*>(base + off_feq[g]);
        hg[g].f_seq = hg[g].f_eq + 256;
        hg[g].bar_off = reinterpret_cast<const uint8_t*>(base + off_bar[g]);
        hg[g].sh_off = hg[g].bar_off + static_cast<size_t>(2) * ((groups[g].n_barcodes + 31) / 32) * 64 * 32;
        hg[g].pol = pol;
        hg[g].ov = reinterpret_cast<const int*>(base + off_ov[g]);
    }
    cudaError_t e1 = cudaMemcpy(T.d_blob.p, blob.data(), blob.size() * 8, cudaMemcpyHostToDevice);
    cudaError_t e2 = cudaMemcpy(T.d_groups.p, hg.data(), sizeof(DevGroup) * n_groups, cudaMemcpyHostToDevice);
    cudaError_t e3 = cudaMemcpy(T.d_code.p, kAlpha.code, 256, cudaMemcpyHostToDevice);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) return (ctx_error(c, "cudaMemcpy of the pattern tables failed"), BB_ERR_CUDA);
    T.host = hg; T.n = n_groups; T.max_trace_cols = max_trace; T.max_nw = max_nw; T.max_region = max_region; T.max_bar_len = max_bar_len; T.max_own_rows = max_own_rows;
    T.max_k = 0; for (int g = 0; g < n_groups; g++) T.max_k = std::max(T.max_k, groups[g].k_flank);
    return BB_OK;
}

static int validate_offsets(bb_ctx* c, const uint64_t* offsets, uint32_t n_reads) {
    for (uint32_t r = 0; r < n_reads; r++) {
        if (offsets[r + 1] < offsets[r]) { ctx_error(c, "offsets must be non-decreasing"); return BB_ERR_INVALID; }
        if (offsets[r + 1] - offsets[r] > bb::kMaxReadLen) { ctx_error(c, "read longer than 2^28 bases"); return BB_ERR_INVALID; }
    }
    return BB_OK;
}

int bb_annotate(bb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, bb_row* rows, uint64_t rows_cap, uint64_t* n_rows) {
    if (!c || !offsets || !n_rows || (n_reads && !bases)) return BB_ERR_INVALID;
    int rc = validate_offsets(c, offsets, n_reads);
    if (rc != BB_OK) return rc;
    bb::Engine& E = c->eng[0];
    rc = E.run_host(bases, offsets, n_reads, n_rows);
    c->last_engine = 0;
    if (rc != BB_OK) { c->err = E.err; return rc; }
    c->total_reads += n_reads; c->kept_reads += E.last_kept;
    if (*n_rows > rows_cap) { ctx_error(c, "row buffer too small"); return BB_ERR_OVERFLOW; }
    if (*n_rows) std::memcpy(rows, E.h_rows, *n_rows * sizeof(bb_row));
    return BB_OK;
}

int bb_annotate_device(bb_ctx* c, const void* d_bases, const void* d_offsets, uint32_t n_reads, uint64_t total_bytes, void* stream, uint64_t* n_rows) {
    if (!c || !n_rows) return BB_ERR_INVALID;
    bb::Engine& E = c->eng[0];
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : E.stream;
    int rc = E.run(static_cast<const uint8_t*>(d_bases), static_cast<const uint64_t*>(d_offsets), n_reads, total_bytes, st, n_rows);
    c->last_engine = 0;
    if (rc != BB_OK) { c->err = E.err; return rc; }
    c->total_reads += n_reads; c->kept_reads += E.last_kept;
    return BB_OK;
}

int bb_fetch_rows(bb_ctx* c, bb_row* rows, uint64_t rows_cap, uint64_t* n_rows) {
    if (!c || !n_rows) return BB_ERR_INVALID;
    bb::Engine& E = c->eng[c->last_engine];
    *n_rows = E.last_rows;
    if (E.last_rows > rows_cap) { ctx_error(c, "row buffer too small"); return BB_ERR_OVERFLOW; }
    if (E.last_rows) {
        cudaSetDevice(E.device);
        cudaError_t e = cudaMemcpy(rows, E.d_rows_out.p, E.last_rows * sizeof(bb_row), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { ctx_error(c, std::string("cudaMemcpy failed: ") + cudaGetErrorString(e)); return BB_ERR_CUDA; }
    }
    return BB_OK;
}

int bb_submit(bb_ctx* c, const uint8_t* bases, const uint64_t* offsets, uint32_t n_reads, uint64_t batch_tag) {
    if (!c || !offsets || (n_reads && !bases)) return BB_ERR_INVALID;
    int rc = validate_offsets(c, offsets, n_reads);
    if (rc != BB_OK) return rc;
    std::unique_lock<std::mutex> lk(c->mu);
    if (!c->workers_started) {
        for (int i = 0; i < BB_MAX_INFLIGHT; i++) c->workers[i] = std::thread(worker_main, c, i);
        c->workers_started = true;
    }
    if (c->order.size() >= BB_MAX_INFLIGHT) { ctx_error(c, "too many batches in flight: call bb_collect first"); return BB_ERR_INVALID; }
    const int idx = static_cast<int>(c->submitted % BB_MAX_INFLIGHT);
    auto* job = new bb_job{bases, offsets, n_reads, batch_tag, idx};
    c->queue[idx].push_back(job);
    c->order.push_back(job);
    c->submitted++;
    lk.unlock();
    c->cv.notify_all();
    return BB_OK;
}

int bb_submit_packed(bb_ctx* c, const uint8_t* crumbs, uint64_t n_bases, const uint64_t* exc, uint64_t n_exc, const uint64_t* offsets,
                     uint32_t n_reads, uint64_t batch_tag) {
    if (!c || !offsets || (n_reads && !crumbs) || (n_exc && !exc)) return BB_ERR_INVALID;
    int rc = validate_offsets(c, offsets, n_reads);
    if (rc != BB_OK) return rc;
    if (offsets[0] != 0 || offsets[n_reads] != n_bases) { ctx_error(c, "offsets must start at 0 and end at n_bases"); return BB_ERR_INVALID; }
    std::unique_lock<std::mutex> lk(c->mu);
    if (!c->workers_started) {
        for (int i = 0; i < BB_MAX_INFLIGHT; i++) c->workers[i] = std::thread(worker_main, c, i);
        c->workers_started = true;
    }
    if (c->order.size() >= BB_MAX_INFLIGHT) { ctx_error(c, "too many batches in flight: call bb_collect first"); return BB_ERR_INVALID; }
    const int idx = static_cast<int>(c->submitted % BB_MAX_INFLIGHT);
    auto* job = new bb_job{crumbs, offsets, n_reads, batch_tag, idx};
    job->packed = true; job->n_bases = n_bases; job->exc = exc; job->n_exc = n_exc;
    c->queue[idx].push_back(job);
    c->order.push_back(job);
    c->submitted++;
    lk.unlock();
    c->cv.notify_all();
    return BB_OK;
}

int bb_reserve(bb_ctx* c, uint32_t max_reads, uint64_t max_bases) {
    if (!c) return BB_ERR_INVALID;
    if (!c->gt.n) { ctx_error(c, "bb_reserve: no query groups set"); return BB_ERR_INVALID; }
    {
        std::unique_lock<std::mutex> lk(c->mu);
        if (!c->order.empty()) { ctx_error(c, "bb_reserve: batches in flight"); return BB_ERR_INVALID; }
        if (!c->workers_started) {
            for (int i = 0; i < BB_MAX_INFLIGHT; i++) c->workers[i] = std::thread(worker_main, c, i);
            c->workers_started = true;
        }
        for (int i = 0; i < BB_MAX_INFLIGHT; i++) {               // one per engine, all at once
            const int idx = static_cast<int>(c->submitted % BB_MAX_INFLIGHT);
            auto* job = new bb_job{nullptr, nullptr, max_reads, 0, idx};
            job->warm = true; job->n_bases = max_bases;
            c->queue[idx].push_back(job);
            c->order.push_back(job);
            c->submitted++;
        }
    }
    c->cv.notify_all();
    int rc = BB_OK;
    std::unique_lock<std::mutex> lk(c->mu);
    while (!c->order.empty()) {
        bb_job* job = c->order.front();
        c->cv.wait(lk, [&] { return job->done; });
        c->order.pop_front();
        if (job->rc != BB_OK && rc == BB_OK) { rc = job->rc; c->err = c->eng[job->engine].err; }
        delete job;
    }
    return rc;
}

int bb_collect(bb_ctx* c, uint64_t* batch_tag, const bb_row** rows, uint64_t* n_rows) {
    if (!c || !rows || !n_rows) return BB_ERR_INVALID;
    std::unique_lock<std::mutex> lk(c->mu);
    if (c->order.empty()) { ctx_error(c, "nothing in flight"); return BB_ERR_INVALID; }
    bb_job* job = c->order.front();
    c->cv.wait(lk, [&] { return job->done; });
    c->order.pop_front();
    const int idx = job->engine, rc = job->rc;
    if (batch_tag) *batch_tag = job->tag;
    *n_rows = job->n_rows;
    *rows = c->eng[idx].h_rows;      // this engine's buffer: rewritten when the engine runs its next batch (BB_MAX_INFLIGHT submits later)
    if (rc == BB_OK) { c->total_reads += job->n_reads; c->kept_reads += c->eng[idx].last_kept; c->last_engine = idx; }
    else c->err = c->eng[idx].err;
    delete job;
    return rc;
}

int bb_counters(const bb_ctx* c, uint64_t out[3]) {
    if (!c || !out) return BB_ERR_INVALID;
    out[0] = c->total_reads; out[1] = c->kept_reads; out[2] = c->total_reads - c->kept_reads;
    return BB_OK;
}

int bb_last_stage_ms(bb_ctx* c, float out[5]) {
    if (!c || !out) return BB_ERR_INVALID;
    for (int i = 0; i < 5; i++) out[i] = c->eng[c->last_engine].stage_ms[i];
    return BB_OK;
}

uint64_t bb_h2d_bytes(const bb_ctx* c) {
    uint64_t n = 0;
    if (c) for (const auto& e : c->eng) n += e.h2d_bytes;
    return n;
}

uint64_t bb_kernel_launches(const bb_ctx* c) {
    uint64_t n = 0;
    if (c) for (const auto& e : c->eng) n += e.launches;
    return n;
}

int bb_fetch_flank_hits(bb_ctx* c, int32_t* out6, uint64_t cap, uint64_t* n_hits) {
    if (!c || !n_hits) return BB_ERR_INVALID;
    bb::Engine& E = c->eng[c->last_engine];
    *n_hits = E.last_hits;
    if (E.last_hits > cap) { ctx_error(c, "hit buffer too small"); return BB_ERR_OVERFLOW; }
    if (E.last_hits == 0) return BB_OK;
    cudaSetDevice(E.device);
    if (E.d_hits6.ensure(static_cast<size_t>(E.last_hits) * 24) != cudaSuccess) { ctx_error(c, "cudaMalloc failed"); return BB_ERR_CUDA; }
    bb::k_export_hits<<<(E.last_hits + 255) / 256, 256, 0, E.stream>>>(E.d_hits.as<bb::Hit>(), E.last_hits, E.d_hits6.as<int32_t>());
    E.launches++;
    cudaError_t e = cudaMemcpyAsync(out6, E.d_hits6.p, static_cast<size_t>(E.last_hits) * 24, cudaMemcpyDeviceToHost, E.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(E.stream);
    if (e != cudaSuccess) { ctx_error(c, std::string("export failed: ") + cudaGetErrorString(e)); return BB_ERR_CUDA; }
    return BB_OK;
}

int bb_pack_nibbles(const uint8_t* src, uint64_t n, uint8_t* dst) {
    if ((!src || !dst) && n) return BB_ERR_INVALID;
    bb::pack_nibbles(src, n, dst, bb::kAlpha.code, bb::pack_default_threads());
    return BB_OK;
}

int bb_pack_crumbs(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t* exc, uint64_t exc_cap, uint64_t* n_exc) {
    if (((!src || !dst) && n) || !n_exc || (!exc && exc_cap)) return BB_ERR_INVALID;
    size_t used = 0; bool over = false;
    bb::pack_crumbs(src, n, dst, exc, exc_cap, &used, &over, bb::kAlpha.code, bb::pack_default_threads());
    *n_exc = used;
    return over ? BB_ERR_OVERFLOW : BB_OK;
}

int bb_pack_crumbs_append(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t* pos, uint64_t* exc, uint64_t exc_cap, uint64_t* n_exc) {
    if ((!src && n) || !dst || !pos || !n_exc || (!exc && exc_cap)) return BB_ERR_INVALID;
    size_t used = static_cast<size_t>(*n_exc);
    const bool ok = bb::crumbs_append(src, static_cast<size_t>(n), dst, pos, exc, static_cast<size_t>(exc_cap), &used, bb::kAlpha.code);
    *n_exc = used;
    return ok ? BB_OK : BB_ERR_OVERFLOW;
}

int bb_pack_crumbs_append_line(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t* pos, uint64_t* exc, uint64_t exc_cap, uint64_t* n_exc,
                               uint64_t* line_len, int* found) {
    if ((!src && n) || !dst || !pos || !n_exc || !line_len || !found || (!exc && exc_cap)) return BB_ERR_INVALID;
    size_t used = static_cast<size_t>(*n_exc), len = 0; bool nl = false;
    const bool ok = bb::crumbs_append_line(src, static_cast<size_t>(n), dst, pos, exc, static_cast<size_t>(exc_cap), &used, bb::kAlpha.code, &len, &nl);
    *n_exc = used; *line_len = len; *found = nl ? 1 : 0;
    return ok ? BB_OK : BB_ERR_OVERFLOW;
}

void* bb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}
void bb_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
