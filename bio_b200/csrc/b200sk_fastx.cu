// Record feeder: the caller side of the sketching path.  What seqio/fastx.Reader.Read + parseRecord
// (seqio/fastx/reader.go:233-471) do one record at a time -- find the records of a FASTA/FASTQ text, strip
// line ends, concatenate each record's sequence lines -- done for a whole chunk of text in HBM, producing
// directly what the sketching kernels take: the packed bases and read_off[n+1], plus per-record text
// offsets (record start, quality start) from which the host shim slices Name/ID/Desc/Qual lazily.
//
// Reference behaviour kept (reader.go):
//   * format from the first byte that is not '\n': '>' FASTA, '@' FASTQ, anything else ErrNotFASTXFormat
//     (:271-304);
//   * a record starts at a delimiter that follows '\n' (:310-325); '>'/'@' elsewhere in a line is data;
//   * FASTA: the sequence is every following line up to the next record, each without its '\n' and one
//     trailing '\r' (:383-393, dropCR :535-541); empty lines contribute nothing;
//   * FASTQ: the line after the header is the sequence, a non-empty line starting with '+' switches to
//     quality (:396-412), sequence and quality lengths must agree (:415-417, ErrUnequalSeqAndQual); an '@'
//     that starts a quality line is not a record start because the record is then not complete (:328-340);
//   * a last record without a final '\n' is complete (:352-364); empty lines after the last record are not
//     a record.
// FASTQ records of exactly four lines (what every sequencer writes) take the parallel path (k_fq_scan); a chunk that
// is not made of such records is parsed by the reference's general rule (k_fq_general: multi-line sequence and
// quality, blank lines between records).  The alphabet is guessed from the first record and every letter checked
// against it (reader.go:430-452), reported per record instead of ending the stream.
//
// Kernels (all memory-bound, one pass each over what they touch):
//   k_fx_detect   first record byte and format
//   k_fx_lines    line-start table L[] in one pass (per-tile newline counts, decoupled look-back); the table is
//                 sized from a bound, the pass also counts the lines and the FASTA record delimiters
//   k_fq_scan     FASTQ: per record validate + sequence length -> read_off / rec_off / qual_off (look-back scan),
//                 then the sequence lines -> packed bases in the same pass
//   k_fa_scan     FASTA: per line header flag + sequence bytes -> out_pos per line, read_off / rec_off per record
//   k_fa_copy     FASTA (and general FASTQ) sequence lines -> packed bases (a warp per 32 lines)
//   k_fq_general  FASTQ of any line structure: one warp walks the line table (fallback of k_fq_scan)
//   k_fx_letters / k_fx_validate   alphabet guess from the first record, per-record validity flags
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>

#include "../../include/b200sketch.h"
#include "b200sk_device.cuh"

namespace b200sk {

// internal hooks into the context (b200sk_api.cu)
int ctx_device(b200sk_ctx *ctx);
cudaStream_t ctx_stream(b200sk_ctx *ctx);
void **ctx_fx_slot(b200sk_ctx *ctx);
void ctx_set_error(b200sk_ctx *ctx, const char *msg);
void ctx_add_launches(b200sk_ctx *ctx, uint64_t n);

namespace {

struct Buf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct FxState {
    Buf meta, lines, outpos, bases, read_off, rec_off, qual_off, state_a, state_b, invalid;
    Buf text;                                   // host path: the chunk in HBM
    Buf o_val, o_pos, o_off, o_status;          // host path: sketch outputs in HBM
    PinBuf h_val, h_pos, h_off, h_status, h_meta; // host path: what the caller reads
};

// meta words (u64): 0 format, 1 start0, 2 newlines, 3 line-initial '>' count, 4 last byte, 5 ticket,
// 6 error record (min), 7 max sequence length, 8 total bases, 9 records (FASTA), 10 flags, 11 ticket 2
// 12 text offset where the records complete in this chunk end (general FASTQ), 13 first record with an invalid letter,
// 16..19 the 256 bits 'byte value occurs in the first record' (alphabet guess)
enum { M_FORMAT = 0, M_START, M_NL, M_HDR, M_LAST, M_TICKET, M_BADREC, M_MAXLEN, M_TOTAL, M_NREC, M_FLAGS, M_TICKET2, M_CONSUMED, M_INVALID, M_ABITS = 16, M_WORDS = 24 };

// ------------------------------------------------------------------ kernels
// reader.go:271-304: the first byte that is not '\n' decides
// limit: how far to look.  While the format is still unknown the reader gives up after 10 241 newlines anyway
// (reader.go:286-294), so 1 MiB is plenty; a chunk that CONTINUES a file (format given) may start with any number of
// blank lines and is searched to its end -- eight bytes per lane and trip beyond the first MiB.
__global__ void k_fx_detect(const uint8_t *__restrict__ t, uint64_t n, unsigned long long *meta, uint64_t limit) {
    const unsigned lane = threadIdx.x;
    uint64_t fmt = 0, start = n;
    uint64_t base = 0;
    for (; base < n && base < limit && base < (1ull << 20); base += 32) {
        const uint64_t i = base + lane;
        const uint8_t b = i < n ? t[i] : (uint8_t)'\n';
        const unsigned other = __ballot_sync(0xffffffffu, b != '\n');
        if (other) {
            const int f = __ffs(other) - 1;
            const uint8_t c = (uint8_t)__shfl_sync(0xffffffffu, (unsigned)b, f);
            start = base + f;
            fmt = c == '>' ? B200SK_FASTX_FASTA : c == '@' ? B200SK_FASTX_FASTQ : 0;
            break;
        }
    }
    if (start == n) {
        for (; base < n && base < limit; base += 256) { // lane l: bytes base + 8 l .. + 7
            uint32_t firstbad = 8;
            uint8_t cb = 0;
#pragma unroll
            for (int j = 7; j >= 0; j--) {
                const uint64_t i = base + 8ull * lane + (uint64_t)j;
                const uint8_t b = i < n ? t[i] : (uint8_t)'\n';
                if (b != '\n') { firstbad = (uint32_t)j; cb = b; }
            }
            const unsigned other = __ballot_sync(0xffffffffu, firstbad < 8);
            if (other) {
                const int f = __ffs(other) - 1;
                const uint32_t j = __shfl_sync(0xffffffffu, firstbad, f);
                const uint8_t c = (uint8_t)__shfl_sync(0xffffffffu, (unsigned)cb, f);
                start = base + 8ull * (uint64_t)f + j;
                fmt = c == '>' ? B200SK_FASTX_FASTA : c == '@' ? B200SK_FASTX_FASTQ : 0;
                break;
            }
        }
    }
    if (lane == 0) {
        meta[M_FORMAT] = fmt;
        meta[M_START] = start;
        meta[M_LAST] = n ? t[n - 1] : '\n';
    }
}

// bit 7 of every byte of w that equals the byte replicated in c4 (exact: no borrow between bytes)
__device__ __forceinline__ uint32_t eq_msb(uint32_t w, uint32_t c4) {
    const uint32_t x = w ^ c4;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}
// bit j of the result: byte j of the 16-byte vector equals c
__device__ __forceinline__ uint32_t eq_mask16(const uint4 v, uint32_t c4) {
    const uint32_t a = __vcmpeq4(v.x, c4) & 0x01010101u, b = __vcmpeq4(v.y, c4) & 0x01010101u;
    const uint32_t c = __vcmpeq4(v.z, c4) & 0x01010101u, d = __vcmpeq4(v.w, c4) & 0x01010101u;
    // gather the four flag bits of a word into a nibble: bits 0,8,16,24 -> 0,1,2,3
    auto nib = [](uint32_t x) { return (x * 0x00204081u) >> 21 & 0xfu; }; // (1 + 2^7 + 2^14 + 2^21) spreads, top nibble collects
    return nib(a) | (nib(b) << 4) | (nib(c) << 8) | (nib(d) << 12);
}
// the same mask from the four eq_msb words of the vector (bit 7 of byte i of word w -> bit 4 w + i)
__device__ __forceinline__ uint32_t nib_of_msb(uint32_t a) { return (((a >> 7) * 0x00204081u) >> 21) & 0xfu; }
__device__ __forceinline__ uint32_t mask16_of_msb(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    return nib_of_msb(a) | (nib_of_msb(b) << 4) | (nib_of_msb(c) << 8) | (nib_of_msb(d) << 12);
}
__device__ __forceinline__ uint32_t range_mask16(uint64_t pos, uint64_t lo, uint64_t hi) { // bytes pos+j in [lo, hi)
    uint32_t m = 0xffffu;
    if (pos < lo) m &= lo - pos >= 16 ? 0u : (0xffffu << (uint32_t)(lo - pos));
    if (pos + 16 > hi) m &= hi <= pos ? 0u : (0xffffu >> (uint32_t)(pos + 16 - hi));
    return m & 0xffffu;
}

// L[i] = start of line i: L[0] = start0, L[i] = position after the i-th newline at or after start0.
// One pass: tiles of 32 KB = 8 rows of 4 KB; in a row thread t owns bytes [16 t, 16 t + 16), so every load
// instruction of a warp is one contiguous 512-byte run (eight such loads in flight per thread).  Newline counts are
// scanned per (row, warp) -- position order -- and the per-tile totals ordered by a decoupled look-back.
#define FX_TILE 32768ull
#define FX_ROW 4096ull
__global__ void __launch_bounds__(256) k_fx_lines(const uint8_t *__restrict__ t, uint64_t n, unsigned long long *meta,
                                                  uint64_t *tile_state, uint64_t *__restrict__ L, uint64_t cap) {
    __shared__ uint32_t ws[64]; // newlines of (row u, warp w) at [8 u + w], then their exclusive prefix
    __shared__ uint64_t s_tile, s_base;
    const uint64_t start0 = meta[M_START];
    const bool fasta = meta[M_FORMAT] == B200SK_FASTX_FASTA;
    const uint64_t ntiles = (n + FX_TILE - 1) / FX_TILE;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (blockIdx.x == 0 && tid == 0 && fasta && start0 < n) atomicAdd(meta + M_HDR, 1ULL); // the first record
    for (;;) {
        if (tid == 0) s_tile = atomicAdd(meta + M_TICKET, 1ULL);
        __syncthreads();
        const uint64_t tile = s_tile;
        if (tile >= ntiles) break;
        const uint64_t pos0 = tile * FX_TILE + (uint64_t)tid * 16;
        uint32_t m[8], inc[8];
        if (tile * FX_TILE >= start0 && (tile + 1) * FX_TILE <= n) { // a tile inside the text: no range checks
            uint4 xs[8];
#pragma unroll
            for (int u = 0; u < 8; u++) xs[u] = reinterpret_cast<const uint4 *>(t)[(pos0 + u * FX_ROW) / 16];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t a = eq_msb(xs[u].x, 0x0a0a0a0au), b = eq_msb(xs[u].y, 0x0a0a0a0au);
                const uint32_t c = eq_msb(xs[u].z, 0x0a0a0a0au), d = eq_msb(xs[u].w, 0x0a0a0a0au);
                m[u] = (a | b | c | d) ? mask16_of_msb(a, b, c, d) : 0u;
                inc[u] = __popc(m[u]);
            }
        } else { // the tile with the first record / the end of the text
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint64_t pos = pos0 + u * FX_ROW;
                m[u] = 0;
                if (pos < n && pos + 16 > start0)
                    m[u] = eq_mask16(reinterpret_cast<const uint4 *>(t)[pos / 16], 0x0a0a0a0au) & range_mask16(pos, start0, n);
                inc[u] = __popc(m[u]);
            }
        }
        // eight independent warp scans (inclusive), one per row
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc[u], o);
                if (lane >= (uint32_t)o) inc[u] += v;
            }
        }
        if (lane == 31) {
#pragma unroll
            for (int u = 0; u < 8; u++) ws[8 * u + warp] = inc[u];
        }
        __syncthreads();
        if (tid < 32) { // 64 (row, warp) counts -> exclusive prefix in position order; tile total -> look-back
            const uint32_t v0 = ws[2 * lane], v1 = ws[2 * lane + 1];
            uint32_t sum = v0 + v1;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, sum, o);
                if (lane >= (uint32_t)o) sum += v;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, sum, 31);
            ws[2 * lane] = sum - v0 - v1;
            ws[2 * lane + 1] = sum - v1;
            const uint64_t b = lookback_exclusive(tile_state, tile, total);
            if (tid == 0) {
                s_base = b;
                if (tile + 1 == ntiles) meta[M_NL] = b + total; // all newlines of the text
            }
        }
        __syncthreads();
        const uint64_t base = s_base;
        uint32_t hd = 0;
#pragma unroll
        for (int u = 0; u < 8; u++) {
            uint32_t mm = m[u];
            uint64_t idx = base + ws[8 * u + warp] + inc[u] - __popc(mm) + 1; // line index this chunk's first newline opens
            while (mm) {
                const int j = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint64_t ls = pos0 + u * FX_ROW + j + 1;
                if (idx < cap) L[idx] = ls; // a table sized from a bound: the host re-runs with the exact size
                idx++;
                if (fasta && ls < n && t[ls] == '>') hd++; // a record delimiter: '>' right after a newline
            }
        }
        if (fasta) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) hd += __shfl_xor_sync(0xffffffffu, hd, o);
            if (lane == 0 && hd) atomicAdd(meta + M_HDR, (unsigned long long)hd);
        }
        __syncthreads();
    }
}
// the two ends of the table
__global__ void k_fx_lines_finish(uint64_t n, const unsigned long long *meta, uint64_t *L, uint64_t cap) {
    const uint64_t start0 = meta[M_START], nl = meta[M_NL];
    L[0] = start0;
    // a last line without '\n': its (virtual) newline sits at n, so the "next line" starts at n + 1
    if (n > start0 && meta[M_LAST] != '\n' && nl + 1 < cap) L[nl + 1] = n + 1;
}

__device__ __forceinline__ uint32_t line_len(const uint8_t *t, uint64_t s, uint64_t next) { // without '\n' and one '\r'
    const uint64_t e = next - 1; // position of the newline
    uint32_t len = (uint32_t)(e - s);
    if (len && t[e - 1] == '\r') len--; // dropCR, reader.go:535-541
    return len;
}

// FASTQ, one thread per record of four lines (reader.go:396-417)
__global__ void __launch_bounds__(256) k_fq_scan(const uint8_t *__restrict__ t, const uint64_t *__restrict__ L,
                                                 uint64_t nrec, uint64_t nlines, int final, unsigned long long *meta,
                                                 uint64_t *tile_state, uint64_t *read_off, uint64_t *rec_off,
                                                 uint64_t *qual_off, uint8_t *__restrict__ bases) {
    __shared__ uint32_t warp_sums[34];
    __shared__ uint64_t s_tile, s_base;
    const uint32_t tid = threadIdx.x;
    const uint64_t ntiles = (nrec + 255) / 256;
    if (blockIdx.x == 0 && tid == 0) {
        // lines after the last whole record: only empty ones are not an error, and only at the end of the text
        for (uint64_t i = 4 * nrec; i < nlines; i++)
            if (!final || line_len(t, L[i], L[i + 1]) != 0) atomicMin(meta + M_BADREC, (unsigned long long)nrec);
        if (nrec == 0) read_off[0] = 0;
    }
    for (;;) {
        if (tid == 0) s_tile = atomicAdd(meta + M_TICKET2, 1ULL);
        __syncthreads();
        const uint64_t tile = s_tile;
        if (tile >= ntiles) break;
        const uint64_t r = tile * 256 + tid;
        uint32_t len = 0;
        uint64_t s = 0;
        if (r < nrec) {
            const uint64_t h = L[4 * r], p = L[4 * r + 2], q = L[4 * r + 3], e = L[4 * r + 4];
            s = L[4 * r + 1];
            len = line_len(t, s, p);
            const uint32_t qlen = line_len(t, q, e);
            // header '@'; a NON-EMPTY line starting with '+' (reader.go:399); equal lengths (:415)
            const bool ok = t[h] == '@' && p + 1 < q && t[p] == '+' && len == qlen;
            if (!ok) {
                atomicMin(meta + M_BADREC, (unsigned long long)r);
                len = 0; // a broken record copies nothing: only records with len == qlen count against the
                         // n_bytes / 2 the base buffer holds (the call fails with BAD_FASTQ anyway)
            }
            rec_off[r] = h;
            qual_off[r] = q;
        }
        uint32_t total;
        const uint32_t excl = block_excl_scan(len, warp_sums, &total);
        uint32_t mx = len;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((tid & 31u) == 0 && mx) atomicMax(meta + M_MAXLEN, (unsigned long long)mx);
        if (tid < 32) {
            const uint64_t b = lookback_exclusive(tile_state, tile, total);
            if (tid == 0) s_base = b;
        }
        __syncthreads();
        const uint64_t dst = s_base + excl;
        if (r < nrec) {
            read_off[r] = dst;
            if (r + 1 == nrec) {
                read_off[nrec] = dst + len;
                meta[M_TOTAL] = dst + len;
            }
        }
        // the sequence lines of the warp's 32 records -> packed bases, record after record
        for (int i = 0; i < 32; i++) {
            const uint64_t si = __shfl_sync(0xffffffffu, s, i), di = __shfl_sync(0xffffffffu, dst, i);
            const uint32_t li = __shfl_sync(0xffffffffu, len, i);
            for (uint32_t j = tid & 31u; j < li; j += 32u) bases[di + j] = t[si + j];
        }
        __syncthreads();
    }
}
// FASTA, one thread per line (reader.go:383-393): header lines open a record, every other line adds its
// bytes to the current record.  Two ordered sums: sequence bytes and header count.
__global__ void __launch_bounds__(256) k_fa_scan(const uint8_t *__restrict__ t, const uint64_t *__restrict__ L,
                                                 uint64_t nlines, unsigned long long *meta, uint64_t *state_len,
                                                 uint64_t *state_hdr, uint64_t *outpos, uint64_t *read_off,
                                                 uint64_t *rec_off) {
    __shared__ uint32_t warp_sums[34];
    __shared__ uint64_t s_tile, s_base_len, s_base_hdr;
    const uint32_t tid = threadIdx.x;
    const uint64_t ntiles = (nlines + 255) / 256;
    if (blockIdx.x == 0 && tid == 0 && nlines == 0) { read_off[0] = 0; meta[M_NREC] = 0; meta[M_TOTAL] = 0; }
    for (;;) {
        if (tid == 0) s_tile = atomicAdd(meta + M_TICKET2, 1ULL);
        __syncthreads();
        const uint64_t tile = s_tile;
        if (tile >= ntiles) break;
        const uint64_t i = tile * 256 + tid;
        uint32_t len = 0, hdr = 0, cr = 0;
        uint64_t s = 0;
        if (i < nlines) {
            s = L[i];
            const uint64_t nx = L[i + 1];
            hdr = (nx - 1 > s && t[s] == '>') ? 1u : 0u; // an empty line is not a header
            const uint32_t full = (uint32_t)(nx - 1 - s);
            len = hdr ? 0u : line_len(t, s, nx);
            cr = hdr ? 0u : full - len;
        }
        uint32_t total_len, total_hdr;
        const uint32_t excl_len = block_excl_scan(len, warp_sums, &total_len);
        const uint32_t excl_hdr = block_excl_scan(hdr, warp_sums, &total_hdr);
        if (tid < 32) {
            const uint64_t b = lookback_exclusive(state_len, tile, total_len);
            const uint64_t c = lookback_exclusive(state_hdr, tile, total_hdr);
            if (tid == 0) { s_base_len = b; s_base_hdr = c; }
        }
        __syncthreads();
        if (i < nlines) {
            const uint64_t op = s_base_len + excl_len;
            outpos[i] = op | ((uint64_t)hdr << 63) | ((uint64_t)cr << 62); // what k_fa_copy needs to know of the line
            if (hdr) {
                const uint64_t r = s_base_hdr + excl_hdr;
                read_off[r] = op;
                rec_off[r] = s;
            }
            if (i + 1 == nlines) {
                const uint64_t nrec = s_base_hdr + excl_hdr + hdr;
                read_off[nrec] = op + len;
                meta[M_NREC] = nrec;
                meta[M_TOTAL] = op + len;
            }
        }
        __syncthreads();
    }
}
// A warp takes 32 consecutive lines: line starts, output positions and the header / CR flags (top bits of outpos,
// left there by k_fa_scan) arrive in three coalesced loads, so no text byte is read before the copy itself and the
// 32 lines' copies are independent of each other.
__global__ void __launch_bounds__(256) k_fa_copy(const uint8_t *__restrict__ t, const uint64_t *__restrict__ L,
                                                 const uint64_t *__restrict__ outpos, uint64_t nlines, uint64_t limit,
                                                 uint8_t *__restrict__ bases) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t nwarp = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t base = ((uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; base < nlines;
         base += nwarp * 32) {
        const uint64_t i = base + lane;
        uint64_t s = 0, d = 0;
        uint32_t len = 0;
        if (i < nlines) {
            s = L[i];
            const uint64_t nx = L[i + 1], op = outpos[i];
            d = op & ((1ull << 62) - 1);
            // not: headers, lines of a record that is not complete in this chunk
            if (!(op >> 63) && s < limit) len = (uint32_t)(nx - 1 - s) - (uint32_t)((op >> 62) & 1u);
        }
        uint32_t todo = __ballot_sync(0xffffffffu, len != 0);
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint64_t sk = __shfl_sync(0xffffffffu, s, k), dk = __shfl_sync(0xffffffffu, d, k);
            const uint32_t lk = __shfl_sync(0xffffffffu, len, k);
            for (uint32_t j = lane; j < lk; j += 32u) bases[dk + j] = t[sk + j];
        }
    }
}

// FASTQ with any line structure (multi-line sequence / quality, blank lines between records): the reference's own
// rule, reader.go:308-345 + parseRecord :396-417 -- a record runs from its header to the next line-initial '@' at
// which the quality is exactly as long as the sequence; an '@' line met earlier belongs to the record (its quality
// is still shorter), one met later is ErrBadFASTQFormat; inside a record everything before the first non-empty '+'
// line is sequence, everything after it quality.  The decision at every '@' line depends on all lines before it, so
// ONE warp walks the line table: 32 lines per trip (starts, lengths and first bytes fetched by the 32 lanes at
// once), then the state machine over those 32 in registers.  It is the fallback of k_fq_scan -- four-line records
// never come here.
// Outputs: out_pos per line in k_fa_copy's format (sequence lines with their place in `bases`, every other line
// flagged not-to-copy), read_off / rec_off / qual_off per record, meta: records, bases, longest sequence, first bad
// record, the text offset where the records that are complete in this chunk end.
__global__ void __launch_bounds__(32) k_fq_general(const uint8_t *__restrict__ t, const uint64_t *__restrict__ L,
                                                   uint64_t nlines, uint64_t n_bytes, int final, unsigned long long *meta,
                                                   uint64_t *outpos, uint64_t *read_off, uint64_t *rec_off,
                                                   uint64_t *qual_off) {
    const uint32_t lane = threadIdx.x;
    // warp-uniform state
    int phase = 0;                 // 0: before a record, 1: sequence lines, 2: quality lines
    uint64_t rec = 0, total = 0;   // records accepted, their bases
    uint64_t p_hdr = 0, p_qual = 0; // the pending record: header line start, first quality line start
    uint64_t seq_len = 0, qual_len = 0, max_len = 0;
    bool bad = false, have_qual = false;
    auto accept = [&]() { // the pending record is complete: its sequence lines were placed at total ..
        if (lane == 0) { read_off[rec] = total; rec_off[rec] = p_hdr; qual_off[rec] = p_qual; }
        max_len = seq_len > max_len ? seq_len : max_len;
        total += seq_len;
        rec++;
    };
    for (uint64_t b = 0; b < nlines && !bad; b += 32) {
        const uint64_t i = b + lane;
        uint64_t s = 0, nx = 0;
        uint32_t len = 0, c0 = 0;
        if (i < nlines) {
            s = L[i];
            nx = L[i + 1];
            len = line_len(t, s, nx);
            c0 = nx - 1 > s ? t[s] : 0u; // first byte of a non-empty line
        }
        uint64_t my_out = 1ull << 63; // not a sequence line
        const uint32_t cnt = (uint32_t)min((uint64_t)32, nlines - b);
        for (uint32_t j = 0; j < cnt; j++) {
            const uint32_t lj = __shfl_sync(0xffffffffu, len, j), cj = __shfl_sync(0xffffffffu, c0, j);
            const uint64_t sj = __shfl_sync(0xffffffffu, s, j), nj = __shfl_sync(0xffffffffu, nx, j);
            if (phase != 0 && cj == '@') { // a candidate record start (reader.go:328-340)
                if (seq_len == qual_len) { accept(); phase = 0; }
                else if (qual_len > seq_len) { bad = true; break; }
                // else the quality is still short: this line belongs to the record
            }
            if (phase == 0) {
                if (cj == '@') { phase = 1; p_hdr = sj; p_qual = 0; have_qual = false; seq_len = 0; qual_len = 0; }
                else if (cj != 0) { bad = true; break; } // only empty lines may come before a record
            } else if (phase == 1) {
                if (lj > 0 && cj == '+') phase = 2; // (:399) quality from the next line on
                else {
                    if (lane == j) my_out = (total + seq_len) | ((uint64_t)((uint32_t)(nj - 1 - sj) - lj) << 62);
                    seq_len += lj;
                }
            } else {
                if (!have_qual) { have_qual = true; p_qual = sj; } // the first quality line (0: the record has none)
                qual_len += lj;
            }
        }
        if (i < nlines) outpos[i] = my_out;
    }
    uint64_t consumed = n_bytes;
    if (!bad && phase != 0) {
        if (final) { // the text ends here: the last record must be complete (reader.go:352-364)
            if (seq_len == qual_len) accept();
            else bad = true;
        } else consumed = p_hdr; // it continues in the next chunk
    }
    if (lane == 0) {
        read_off[rec] = total;
        meta[M_NREC] = rec;
        meta[M_TOTAL] = total;
        meta[M_MAXLEN] = max_len;
        meta[M_BADREC] = bad ? rec : ~0ULL;
        meta[M_CONSUMED] = consumed;
    }
}

// ------------------------------------------------------------------ alphabet (reader.go:430-452)
// The reference guesses the alphabet from the first record (seq.GuessAlphabetLessConservatively over its first
// 10 000 letters, seq/alphabet.go:411-452) and then checks every letter of every record against it
// (Alphabet.IsValid, :234-300; on by default, seq/seq.go:37).  Here: the set of byte values that occur in the first
// record's prefix, and a pass over the packed bases that flags the records holding a letter outside the set given.
__global__ void __launch_bounds__(256) k_fx_letters(const uint8_t *__restrict__ bases, uint64_t n, unsigned long long *bits) {
    __shared__ unsigned int w[8];
    if (threadIdx.x < 8) w[threadIdx.x] = 0;
    __syncthreads();
    for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t b = bases[i];
        atomicOr(&w[b >> 5], 1u << (b & 31u));
    }
    __syncthreads();
    if (threadIdx.x < 4) bits[threadIdx.x] = (unsigned long long)w[2 * threadIdx.x] | ((unsigned long long)w[2 * threadIdx.x + 1] << 32);
}
__global__ void __launch_bounds__(256) k_fx_validate(const uint8_t *__restrict__ bases, uint64_t n,
                                                     const uint64_t *__restrict__ read_off, uint64_t nrec,
                                                     unsigned long long v0, unsigned long long v1, unsigned long long v2,
                                                     unsigned long long v3, uint8_t *__restrict__ invalid,
                                                     unsigned long long *first_invalid) {
    __shared__ uint8_t ok[256];
    {
        const uint32_t b = threadIdx.x;
        const unsigned long long word = b < 64 ? v0 : b < 128 ? v1 : b < 192 ? v2 : v3;
        ok[b] = (uint8_t)((word >> (b & 63u)) & 1ull);
    }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 16;
    for (uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i0 < n; i0 += stride) {
        uint32_t badmask = 0;
        if (i0 + 16 <= n) {
            const uint4 v = *reinterpret_cast<const uint4 *>(bases + i0);
            const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 16; q++) badmask |= (ok[(ws[q >> 2] >> (8 * (q & 3))) & 255u] ? 0u : 1u) << q;
        } else {
            for (uint32_t q = 0; i0 + q < n; q++) badmask |= (ok[bases[i0 + q]] ? 0u : 1u) << q;
        }
        while (badmask) { // rare: find the record of every offending byte
            const uint32_t q = __ffs(badmask) - 1;
            badmask &= badmask - 1;
            const uint64_t at = i0 + q;
            uint64_t lo = 0, hi = nrec; // largest r with read_off[r] <= at
            while (hi - lo > 1) {
                const uint64_t mid = (lo + hi) >> 1;
                if (read_off[mid] <= at) lo = mid; else hi = mid;
            }
            invalid[lo] = 1;
            atomicMin(first_invalid, (unsigned long long)lo);
        }
    }
}

int fail(b200sk_ctx *ctx, cudaError_t e, const char *what) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    ctx_set_error(ctx, buf);
    return B200SK_ERR_CUDA;
}
#define FCK(call)                                          \
    do {                                                   \
        cudaError_t _e = (call);                           \
        if (_e != cudaSuccess) return fail(ctx, _e, #call); \
    } while (0)

// seq/alphabet.go:353-399: all letters (letters + gap + ambiguous) of the reference's alphabets, as 256-bit sets
struct Bits256 { unsigned long long w[4]; };
Bits256 set_of(const char *letters) {
    Bits256 b = {{0, 0, 0, 0}};
    for (const unsigned char *c = (const unsigned char *)letters; *c; c++) b.w[*c >> 6] |= 1ull << (*c & 63);
    return b;
}
bool subset(const Bits256 &a, const Bits256 &b) {
    for (int i = 0; i < 4; i++)
        if (a.w[i] & ~b.w[i]) return false;
    return true;
}
Bits256 letters_of(int alphabet) {
    switch (alphabet) {
    case B200SK_ALPHABET_DNA: return set_of("acgtACGT -.nN");
    case B200SK_ALPHABET_DNA_REDUNDANT: return set_of("acgtryswkmbdhvACGTRYSWKMBDHV -.nN");
    case B200SK_ALPHABET_RNA: return set_of("acguACGU -.nN");
    case B200SK_ALPHABET_RNA_REDUNDANT: return set_of("acguryswkmbdhvACGURYSWKMBDHV -.nN");
    case B200SK_ALPHABET_PROTEIN: return set_of("abcdefghijklmnopqrstuvwyzABCDEFGHIJKLMNOPQRSTUVWYZ -xX*_.");
    default: { Bits256 all = {{~0ull, ~0ull, ~0ull, ~0ull}}; return all; }
    }
}
// seq.GuessAlphabetLessConservatively (seq/alphabet.go:411-452): first match in the order DNA, RNA, DNAredundant,
// RNAredundant, Protein, else Unlimit; DNA and RNA are widened to their redundant forms; no letters: Unlimit.
int guess_alphabet(const Bits256 &present) {
    if (!(present.w[0] | present.w[1] | present.w[2] | present.w[3])) return B200SK_ALPHABET_UNLIMIT;
    if (subset(present, letters_of(B200SK_ALPHABET_DNA))) return B200SK_ALPHABET_DNA_REDUNDANT;
    if (subset(present, letters_of(B200SK_ALPHABET_RNA))) return B200SK_ALPHABET_RNA_REDUNDANT;
    if (subset(present, letters_of(B200SK_ALPHABET_DNA_REDUNDANT))) return B200SK_ALPHABET_DNA_REDUNDANT;
    if (subset(present, letters_of(B200SK_ALPHABET_RNA_REDUNDANT))) return B200SK_ALPHABET_RNA_REDUNDANT;
    if (subset(present, letters_of(B200SK_ALPHABET_PROTEIN))) return B200SK_ALPHABET_PROTEIN;
    return B200SK_ALPHABET_UNLIMIT;
}

FxState *state_of(b200sk_ctx *ctx) {
    void **slot = ctx_fx_slot(ctx);
    if (!*slot) *slot = new FxState();
    return static_cast<FxState *>(*slot);
}

} // namespace

void fx_free(void *p) {
    FxState *s = static_cast<FxState *>(p);
    if (!s) return;
    for (Buf *b : {&s->meta, &s->lines, &s->outpos, &s->bases, &s->read_off, &s->rec_off, &s->qual_off, &s->state_a,
                   &s->state_b, &s->invalid, &s->text, &s->o_val, &s->o_pos, &s->o_off, &s->o_status})
        b->release();
    for (PinBuf *b : {&s->h_val, &s->h_pos, &s->h_off, &s->h_status, &s->h_meta}) b->release();
    delete s;
}

} // namespace b200sk

using namespace b200sk;

extern "C" {

int b200sk_fastx_parse_device(b200sk_ctx *ctx, const uint8_t *d_text, uint64_t n_bytes, int format, int final,
                              void *stream, b200sk_fastx_info *info) {
    if (!ctx || !info || (n_bytes && !d_text)) return B200SK_ERR_BAD_ARG;
    if (((uintptr_t)d_text & 15u) != 0) return B200SK_ERR_BAD_ARG;
    const int alpha_in = (format >> 8) & 0xff; // a chunk that continues a file: 1 + the alphabet guessed from its first record
    format &= 0xff;
    if (format != 0 && format != B200SK_FASTX_FASTA && format != B200SK_FASTX_FASTQ) return B200SK_ERR_BAD_ARG;
    if (alpha_in > B200SK_ALPHABET_PROTEIN + 1) return B200SK_ERR_BAD_ARG;
    memset(info, 0, sizeof(*info));
    info->first_invalid = ~0ULL;
    info->alphabet = alpha_in ? alpha_in - 1 : B200SK_ALPHABET_UNLIMIT;
    FCK(cudaSetDevice(ctx_device(ctx)));
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx_stream(ctx);
    FxState *fx = state_of(ctx);
    FCK(fx->meta.reserve(M_WORDS * 8));
    unsigned long long *meta = (unsigned long long *)fx->meta.p;
    unsigned long long hm[M_WORDS];
    FCK(cudaMemsetAsync(meta, 0, M_WORDS * 8, st));
    if (n_bytes == 0) {
        info->format = format;
        return 0;
    }
    k_fx_detect<<<1, 32, 0, st>>>(d_text, n_bytes, meta, (format & 0xff) ? n_bytes : (1ull << 20));
    ctx_add_launches(ctx, 1);
    // The line table is sized from a bound (a line per 32 bytes of text) and filled in the same pass that
    // counts the lines; only a text with shorter lines on average pays a second pass with the exact size.
    const uint64_t ntiles_l = (n_bytes + FX_TILE - 1) / FX_TILE;
    FCK(fx->state_a.reserve((ntiles_l + 1) * 8));
    uint64_t cap = n_bytes / 32 + 4096;
    for (int attempt = 0;; attempt++) {
        FCK(fx->lines.reserve(cap * 8));
        FCK(cudaMemsetAsync(fx->state_a.p, 0, (ntiles_l + 1) * 8, st));
        FCK(cudaMemsetAsync(meta + M_NL, 0, 16, st));     // newline and header counts
        FCK(cudaMemsetAsync(meta + M_TICKET, 0, 8, st));
        const unsigned blocks = (unsigned)std::min<uint64_t>(ntiles_l, 148ull * 8);
        k_fx_lines<<<blocks, 256, 0, st>>>(d_text, n_bytes, meta, (uint64_t *)fx->state_a.p, (uint64_t *)fx->lines.p, cap);
        k_fx_lines_finish<<<1, 1, 0, st>>>(n_bytes, meta, (uint64_t *)fx->lines.p, cap);
        ctx_add_launches(ctx, 2);
        FCK(cudaMemcpyAsync(hm, meta, M_WORDS * 8, cudaMemcpyDeviceToHost, st));
        FCK(cudaStreamSynchronize(st));
        if (hm[M_NL] + 3 <= cap || attempt) break;
        cap = hm[M_NL] + 3;
    }
    const int detected = (int)hm[M_FORMAT];
    // reader.go:286-294: while the format is still being detected the reader tolerates leading blank lines only up
    // to byte 10240 (more than 100 of them): a '\n' at an index > 10240 with nothing but newlines before it is
    // ErrNotFASTXFormat
    if (format == 0 && std::min<uint64_t>(hm[M_START], n_bytes) >= 10242) {
        info->status = B200SK_ERR_NOT_FASTX;
        return B200SK_ERR_NOT_FASTX;
    }
    if (hm[M_START] >= n_bytes) { // only newlines: nothing to parse
        info->format = format;
        info->consumed = final ? n_bytes : 0;
        return 0;
    }
    if (detected == 0 || (format != 0 && format != detected)) {
        info->status = B200SK_ERR_NOT_FASTX;
        return B200SK_ERR_NOT_FASTX;
    }
    info->format = detected;
    const uint64_t nl = hm[M_NL];
    const bool tail = hm[M_LAST] != '\n'; // a last line without its newline
    // lines the parse may use: without `final` an unterminated last line is not complete
    const uint64_t nlines = nl + ((tail && final) ? 1 : 0);
    uint64_t *L = (uint64_t *)fx->lines.p;
    uint64_t nrec = 0;
    if (detected == B200SK_FASTX_FASTQ) {
        nrec = nlines / 4;
        FCK(fx->read_off.reserve((nrec + 2) * 8));
        FCK(fx->rec_off.reserve((nrec + 2) * 8));
        FCK(fx->qual_off.reserve((nrec + 2) * 8));
        const uint64_t ntiles = (nrec + 255) / 256;
        FCK(fx->state_b.reserve((ntiles + 1) * 8));
        FCK(cudaMemsetAsync(fx->state_b.p, 0, (ntiles + 1) * 8, st));
        const unsigned long long none = ~0ULL;
        FCK(cudaMemcpyAsync(meta + M_BADREC, &none, 8, cudaMemcpyHostToDevice, st));
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(ntiles, 148ull * 8));
        // without `final` the lines after the last whole record simply stay for the next chunk
        FCK(fx->bases.reserve(n_bytes / 2 + 64)); // sequence and quality are equally long: at most half the text
        k_fq_scan<<<blocks, 256, 0, st>>>(d_text, L, nrec, final ? nlines : 4 * nrec, final, meta,
                                          (uint64_t *)fx->state_b.p, (uint64_t *)fx->read_off.p,
                                          (uint64_t *)fx->rec_off.p, (uint64_t *)fx->qual_off.p, (uint8_t *)fx->bases.p);
        ctx_add_launches(ctx, 1);
        FCK(cudaMemcpyAsync(hm, meta, M_WORDS * 8, cudaMemcpyDeviceToHost, st));
        FCK(cudaStreamSynchronize(st));
        if (hm[M_BADREC] != ~0ULL) {
            // not four lines per record: the reference's general rule (multi-line records, blank lines between
            // records), one warp over the line table, then the FASTA copy kernel for the sequence lines
            FCK(fx->read_off.reserve((nlines + 2) * 8));
            FCK(fx->rec_off.reserve((nlines + 2) * 8));
            FCK(fx->qual_off.reserve((nlines + 2) * 8));
            FCK(fx->outpos.reserve((nlines + 2) * 8));
            k_fq_general<<<1, 32, 0, st>>>(d_text, L, nlines, n_bytes, final, meta, (uint64_t *)fx->outpos.p,
                                           (uint64_t *)fx->read_off.p, (uint64_t *)fx->rec_off.p,
                                           (uint64_t *)fx->qual_off.p);
            ctx_add_launches(ctx, 1);
            FCK(cudaMemcpyAsync(hm, meta, M_WORDS * 8, cudaMemcpyDeviceToHost, st));
            FCK(cudaStreamSynchronize(st));
            if (hm[M_BADREC] != ~0ULL) {
                info->status = B200SK_ERR_BAD_FASTQ;
                info->bad_record = hm[M_BADREC];
                return B200SK_ERR_BAD_FASTQ;
            }
            nrec = hm[M_NREC];
            const uint64_t total = hm[M_TOTAL];
            FCK(fx->bases.reserve(total + 64));
            if (nlines && total) {
                const unsigned cb = (unsigned)std::min<uint64_t>((nlines + 255) / 256, 148ull * 16);
                // lines of the record that continues in the next chunk lie at or behind `consumed`: not copied
                k_fa_copy<<<cb, 256, 0, st>>>(d_text, L, (const uint64_t *)fx->outpos.p, nlines,
                                              final ? n_bytes + 2 : hm[M_CONSUMED], (uint8_t *)fx->bases.p);
                ctx_add_launches(ctx, 1);
            }
            info->n_records = nrec;
            info->n_bases = total;
            info->max_read_len = (uint32_t)hm[M_MAXLEN];
            info->consumed = final ? n_bytes : std::min<uint64_t>(hm[M_CONSUMED], n_bytes);
        } else {
            const uint64_t total = nrec ? hm[M_TOTAL] : 0;
            info->n_records = nrec;
            info->n_bases = total;
            info->max_read_len = (uint32_t)hm[M_MAXLEN];
            if (final) info->consumed = n_bytes;
            else {
                uint64_t c = 0;
                if (nrec) {
                    FCK(cudaMemcpyAsync(&c, L + 4 * nrec, 8, cudaMemcpyDeviceToHost, st));
                    FCK(cudaStreamSynchronize(st));
                } else c = hm[M_START];
                info->consumed = std::min<uint64_t>(c, n_bytes);
            }
        }
        info->d_qual_off = (uint64_t *)fx->qual_off.p;
    } else {
        const uint64_t maxrec = hm[M_HDR];
        FCK(fx->read_off.reserve((maxrec + 2) * 8));
        FCK(fx->rec_off.reserve((maxrec + 2) * 8));
        FCK(fx->outpos.reserve((nlines + 2) * 8));
        const uint64_t ntiles = (nlines + 255) / 256;
        FCK(fx->state_a.reserve((ntiles + 1) * 8));
        FCK(fx->state_b.reserve((ntiles + 1) * 8));
        FCK(cudaMemsetAsync(fx->state_a.p, 0, (ntiles + 1) * 8, st));
        FCK(cudaMemsetAsync(fx->state_b.p, 0, (ntiles + 1) * 8, st));
        const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(ntiles, 148ull * 8));
        k_fa_scan<<<blocks, 256, 0, st>>>(d_text, L, nlines, meta, (uint64_t *)fx->state_a.p,
                                          (uint64_t *)fx->state_b.p, (uint64_t *)fx->outpos.p,
                                          (uint64_t *)fx->read_off.p, (uint64_t *)fx->rec_off.p);
        ctx_add_launches(ctx, 1);
        FCK(cudaMemcpyAsync(hm, meta, M_WORDS * 8, cudaMemcpyDeviceToHost, st));
        FCK(cudaStreamSynchronize(st));
        uint64_t found = hm[M_NREC];
        // without `final` the last record may continue in the next chunk: it stays
        nrec = final ? found : (found ? found - 1 : 0);
        uint64_t ro[2] = {0, 0};
        uint64_t limit = n_bytes + 2, total = hm[M_TOTAL];
        if (!final) {
            if (found) {
                FCK(cudaMemcpyAsync(&ro[0], (uint64_t *)fx->read_off.p + nrec, 8, cudaMemcpyDeviceToHost, st));
                FCK(cudaMemcpyAsync(&ro[1], (uint64_t *)fx->rec_off.p + nrec, 8, cudaMemcpyDeviceToHost, st));
                FCK(cudaStreamSynchronize(st));
                total = ro[0];
                limit = ro[1];
            } else { total = 0; limit = hm[M_START]; }
        }
        FCK(fx->bases.reserve(total + 64));
        if (nlines && total) {
            const unsigned cb = (unsigned)std::min<uint64_t>((nlines + 255) / 256, 148ull * 16);
            k_fa_copy<<<cb, 256, 0, st>>>(d_text, L, (const uint64_t *)fx->outpos.p, nlines, limit, (uint8_t *)fx->bases.p);
            ctx_add_launches(ctx, 1);
        }
        info->n_records = nrec;
        info->n_bases = total;
        info->max_read_len = 0; // unknown: the sketching entry points measure it
        info->consumed = final ? n_bytes : std::min<uint64_t>(limit, n_bytes);
        info->d_qual_off = nullptr;
    }
    FCK(cudaMemcpyAsync((uint64_t *)fx->rec_off.p + info->n_records, &info->consumed, 8, cudaMemcpyHostToDevice, st));
    // reader.go:430-452: alphabet of the file (guessed from its first record unless the caller passes it on from an
    // earlier chunk), then every letter of every record checked against it
    if (info->n_records) {
        if (!alpha_in) {
            uint64_t r1 = 0;
            FCK(cudaMemcpyAsync(&r1, (uint64_t *)fx->read_off.p + 1, 8, cudaMemcpyDeviceToHost, st));
            FCK(cudaStreamSynchronize(st));
            k_fx_letters<<<1, 256, 0, st>>>((const uint8_t *)fx->bases.p, std::min<uint64_t>(r1, 10000), meta + M_ABITS);
            ctx_add_launches(ctx, 1);
            Bits256 present;
            FCK(cudaMemcpyAsync(present.w, meta + M_ABITS, 32, cudaMemcpyDeviceToHost, st));
            FCK(cudaStreamSynchronize(st));
            // (a byte >= 0x80 makes the reference's ASCII set construction give up: no alphabet fits -> Unlimit)
            info->alphabet = (present.w[2] | present.w[3]) ? B200SK_ALPHABET_UNLIMIT : guess_alphabet(present);
        }
        FCK(fx->invalid.reserve(info->n_records + 16));
        FCK(cudaMemsetAsync(fx->invalid.p, 0, info->n_records, st));
        if (info->alphabet != B200SK_ALPHABET_UNLIMIT && info->n_bases) {
            const unsigned long long none = ~0ULL;
            FCK(cudaMemcpyAsync(meta + M_INVALID, &none, 8, cudaMemcpyHostToDevice, st));
            const Bits256 ok = letters_of(info->alphabet);
            const unsigned vb = (unsigned)std::min<uint64_t>((info->n_bases / 16 + 255) / 256 + 1, 148ull * 8);
            k_fx_validate<<<vb, 256, 0, st>>>((const uint8_t *)fx->bases.p, info->n_bases, (const uint64_t *)fx->read_off.p,
                                              info->n_records, ok.w[0], ok.w[1], ok.w[2], ok.w[3], (uint8_t *)fx->invalid.p,
                                              meta + M_INVALID);
            ctx_add_launches(ctx, 1);
            FCK(cudaMemcpyAsync(&info->first_invalid, meta + M_INVALID, 8, cudaMemcpyDeviceToHost, st));
        }
        info->d_invalid = (uint8_t *)fx->invalid.p;
    }
    FCK(cudaStreamSynchronize(st));
    info->n_lines = nlines;
    info->d_bases = (uint8_t *)fx->bases.p;
    info->d_read_off = (uint64_t *)fx->read_off.p;
    info->d_rec_off = (uint64_t *)fx->rec_off.p;
    info->d_line_off = L;
    info->status = B200SK_OK;
    return 0;
}

int b200sk_copy_to_host(b200sk_ctx *ctx, void *dst, const void *d_src, uint64_t bytes) {
    if (!ctx || (bytes && (!dst || !d_src))) return B200SK_ERR_BAD_ARG;
    FCK(cudaSetDevice(ctx_device(ctx)));
    FCK(cudaDeviceSynchronize());
    if (bytes) FCK(cudaMemcpy(dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

} // extern "C"

namespace b200sk {

// b200sk_run_fastx in two stages, so that the pipelined reader below can start the next chunk's copy as soon as
// this chunk's parse has told where the next chunk begins.
// stage 1: text chunk (host) -> HBM -> records
static int fx_stage_parse(b200sk_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int format, int final,
                          b200sk_fastx_info *info) {
    FCK(cudaSetDevice(ctx_device(ctx)));
    cudaStream_t st = ctx_stream(ctx);
    FxState *fx = state_of(ctx);
    FCK(fx->text.reserve(((n_bytes + 15) & ~15ull) + 16));
    if (n_bytes) FCK(cudaMemcpyAsync(fx->text.p, text, n_bytes, cudaMemcpyHostToDevice, st));
    return b200sk_fastx_parse_device(ctx, (const uint8_t *)fx->text.p, n_bytes, format, final, st, info);
}
// stage 2: records -> sketches -> pinned host arrays
static int fx_sketch_fetch(b200sk_ctx *ctx, const b200sk_params *p, const b200sk_fastx_info *info, uint64_t **out_val,
                           uint32_t **out_pos, uint64_t **out_off, int32_t **read_status, uint64_t *n_out) {
    FCK(cudaSetDevice(ctx_device(ctx)));
    cudaStream_t st = ctx_stream(ctx);
    FxState *fx = state_of(ctx);
    int rc = 0;
    const uint64_t nrec = info->n_records;
    b200sk_params q = *p;
    if (q.max_read_len == 0) q.max_read_len = info->max_read_len; // FASTQ: measured by the parse
    const uint32_t pw = q.pos_width == 1 ? 1u : q.pos_width == 2 ? 2u : 4u;
    uint64_t cap = b200sk_output_bound(&q, info->n_bases, nrec, 0);
    uint64_t got = 0;
    FCK(fx->o_off.reserve((nrec + 1) * 8));
    FCK(fx->o_status.reserve((nrec + 1) * 4));
    if (nrec == 0) FCK(cudaMemsetAsync(fx->o_off.p, 0, 8, st)); // a chunk without a complete record
    for (int attempt = 0; nrec && attempt < 2; attempt++) {
        FCK(fx->o_val.reserve(cap * 8 + 64));
        if (q.want_pos) FCK(fx->o_pos.reserve(cap * pw + 64));
        rc = b200sk_run_device(ctx, &q, info->d_bases, info->d_read_off, nrec, info->n_bases, (uint64_t *)fx->o_val.p,
                               q.want_pos ? (uint32_t *)fx->o_pos.p : nullptr, (uint64_t *)fx->o_off.p,
                               (int32_t *)fx->o_status.p, cap, st, &got);
        if (rc != B200SK_ERR_CAPACITY) break;
        cap = got;
    }
    if (rc) return rc;
    FCK(fx->h_val.reserve(got * 8 + 8));
    FCK(fx->h_off.reserve((nrec + 1) * 8));
    FCK(fx->h_status.reserve((nrec + 1) * 4));
    if (q.want_pos) FCK(fx->h_pos.reserve(got * pw + 8));
    if (got) FCK(cudaMemcpyAsync(fx->h_val.p, fx->o_val.p, got * 8, cudaMemcpyDeviceToHost, st));
    if (got && q.want_pos) FCK(cudaMemcpyAsync(fx->h_pos.p, fx->o_pos.p, got * pw, cudaMemcpyDeviceToHost, st));
    FCK(cudaMemcpyAsync(fx->h_off.p, fx->o_off.p, (nrec + 1) * 8, cudaMemcpyDeviceToHost, st));
    if (nrec) FCK(cudaMemcpyAsync(fx->h_status.p, fx->o_status.p, nrec * 4, cudaMemcpyDeviceToHost, st));
    FCK(cudaStreamSynchronize(st));
    if (out_val) *out_val = (uint64_t *)fx->h_val.p;
    if (out_pos) *out_pos = q.want_pos ? (uint32_t *)fx->h_pos.p : nullptr;
    if (out_off) *out_off = (uint64_t *)fx->h_off.p;
    if (read_status) *read_status = (int32_t *)fx->h_status.p;
    if (n_out) *n_out = got;
    return 0;
}

} // namespace b200sk

// ------------------------------------------------------------------ pipelined reader
// A whole FASTA/FASTQ text in host memory, handed back chunk by chunk like a reader loop.  Two slots, each
// with its own b200sk_ctx (stream, device and pinned buffers) and its own worker thread; chunk j runs on slot
// j mod 2.  The only dependency between chunks is where the next one starts (the first byte the records of
// this one do not cover), known after stage 1; so chunk j+1's copy and parse overlap chunk j's sketching
// and its copy back, and both overlap the caller consuming chunk j-1.
struct b200sk_fxstream {
    struct Slot {
        b200sk_ctx *ctx = nullptr;
        std::thread th;
        int state = 0; // 0 free, 1 busy, 2 ready
        uint64_t chunk = ~0ull;
        int rc = 0;
        b200sk_fastx_info info;
        uint64_t *val = nullptr, *off = nullptr, n_out = 0;
        uint32_t *pos = nullptr;
        int32_t *status = nullptr;
    } slot[2];
    std::mutex mu;
    std::condition_variable cv;
    b200sk_params params;
    const uint8_t *text = nullptr;
    uint64_t n_bytes = 0, chunk_bytes = 0;
    int format = 0;
    // assignment state (under mu)
    uint64_t next_start = 0, next_index = 0; // the next chunk to hand to a worker
    bool start_ready = false;                // next_start is known (stage 1 of the chunk before has ended)
    bool exhausted = true;                   // no further chunk will be assigned (end of text or error)
    bool stop = false;
    uint64_t deliver = 0;                    // index of the chunk the next call returns
    bool holding = false;                    // the caller still owns the arrays of chunk deliver-1
    bool failed = false;                     // an error has been returned: nothing follows it

    void work(int x) {
        Slot &s = slot[x];
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return stop || (!exhausted && start_ready && s.state == 0 && (next_index & 1) == (uint64_t)x); });
            if (stop) return;
            const uint64_t j = next_index, start = next_start;
            start_ready = false;
            s.state = 1;
            s.chunk = j;
            const int fmt = format;
            lk.unlock();
            uint64_t n = std::min(chunk_bytes, n_bytes - start);
            int rc, final;
            for (;;) { // a chunk that holds no complete record grows until it does
                final = start + n == n_bytes;
                rc = b200sk::fx_stage_parse(s.ctx, text + start, n, fmt, final, &s.info);
                if (rc || final || s.info.consumed) break;
                n = std::min(n * 2, n_bytes - start);
            }
            lk.lock();
            if (rc || final) exhausted = true;
            else {
                // later chunks continue the same file: its format, and the alphabet its first record decided
                if (!(format & 0xff)) format = (format & ~0xff) | s.info.format;
                if (!(format >> 8)) format |= (s.info.alphabet + 1) << 8;
                next_start = start + s.info.consumed;
                next_index = j + 1;
                start_ready = true;
            }
            cv.notify_all();
            lk.unlock();
            if (!rc) rc = b200sk::fx_sketch_fetch(s.ctx, &params, &s.info, &s.val, &s.pos, &s.off, &s.status, &s.n_out);
            lk.lock();
            s.rc = rc;
            if (rc) exhausted = true;
            s.state = 2;
            cv.notify_all();
        }
    }
};

extern "C" {

int b200sk_run_fastx(b200sk_ctx *ctx, const b200sk_params *p, const uint8_t *text, uint64_t n_bytes, int format,
                     int final, b200sk_fastx_info *info, uint64_t **out_val, uint32_t **out_pos, uint64_t **out_off,
                     int32_t **read_status, uint64_t *n_out) {
    if (!ctx || !p || !info || (n_bytes && !text)) return B200SK_ERR_BAD_ARG;
    int rc = b200sk_check_params(p);
    if (rc) return rc;
    if ((rc = b200sk::fx_stage_parse(ctx, text, n_bytes, format, final, info))) return rc;
    return b200sk::fx_sketch_fetch(ctx, p, info, out_val, out_pos, out_off, read_status, n_out);
}

int b200sk_fxstream_rewind(b200sk_fxstream *s, const uint8_t *text, uint64_t n_bytes, int format) {
    if (!s || (n_bytes && !text)) return B200SK_ERR_BAD_ARG;
    std::unique_lock<std::mutex> lk(s->mu);
    s->exhausted = true; // nothing new starts; wait for the chunks in flight
    s->cv.wait(lk, [&] { return s->slot[0].state != 1 && s->slot[1].state != 1; });
    for (auto &sl : s->slot) { sl.state = 0; sl.chunk = ~0ull; }
    s->text = text;
    s->n_bytes = n_bytes;
    s->format = format;
    s->next_start = 0;
    s->next_index = 0;
    s->deliver = 0;
    s->holding = false;
    s->failed = false;
    s->start_ready = true;
    s->exhausted = false;
    s->cv.notify_all();
    return 0;
}

int b200sk_fxstream_open(b200sk_fxstream **out, int device, const b200sk_params *p, const uint8_t *text,
                         uint64_t n_bytes, int format, uint64_t chunk_bytes) {
    if (!out || !p || (n_bytes && !text)) return B200SK_ERR_BAD_ARG;
    *out = nullptr;
    int rc = b200sk_check_params(p);
    if (rc) return rc;
    b200sk_fxstream *s = new b200sk_fxstream();
    s->params = *p;
    s->chunk_bytes = chunk_bytes ? std::max<uint64_t>(chunk_bytes, 64) : (256ull << 20);
    for (auto &sl : s->slot)
        if ((rc = b200sk_create(&sl.ctx, device))) {
            for (auto &t : s->slot) b200sk_destroy(t.ctx);
            delete s;
            return rc;
        }
    for (int x = 0; x < 2; x++) s->slot[x].th = std::thread([s, x] { s->work(x); });
    b200sk_fxstream_rewind(s, text, n_bytes, format);
    *out = s;
    return 0;
}

int b200sk_fxstream_next(b200sk_fxstream *s, b200sk_fastx_info *info, uint64_t **out_val, uint32_t **out_pos,
                         uint64_t **out_off, int32_t **read_status, uint64_t *n_out) {
    if (!s || !info) return B200SK_ERR_BAD_ARG;
    std::unique_lock<std::mutex> lk(s->mu);
    if (s->holding) { // the arrays of the chunk returned last go back to their slot
        b200sk_fxstream::Slot &prev = s->slot[(s->deliver - 1) & 1];
        prev.state = 0;
        s->holding = false;
        s->cv.notify_all();
    }
    if (s->failed) return B200SK_FXSTREAM_END;
    b200sk_fxstream::Slot &sl = s->slot[s->deliver & 1];
    s->cv.wait(lk, [&] {
        return (sl.state == 2 && sl.chunk == s->deliver) || (s->exhausted && sl.state != 1 && sl.chunk != s->deliver);
    });
    if (!(sl.state == 2 && sl.chunk == s->deliver)) return B200SK_FXSTREAM_END;
    s->deliver++;
    s->holding = true;
    *info = sl.info;
    if (sl.rc) {
        s->failed = true;
        return sl.rc;
    }
    if (out_val) *out_val = sl.val;
    if (out_pos) *out_pos = sl.pos;
    if (out_off) *out_off = sl.off;
    if (read_status) *read_status = sl.status;
    if (n_out) *n_out = sl.n_out;
    return 0;
}

uint64_t b200sk_fxstream_kernel_launches(const b200sk_fxstream *s) {
    return s ? b200sk_kernel_launches(s->slot[0].ctx) + b200sk_kernel_launches(s->slot[1].ctx) : 0;
}

const char *b200sk_fxstream_last_error(const b200sk_fxstream *s) {
    if (!s) return "";
    const char *e = b200sk_last_error(s->slot[0].ctx);
    return e && *e ? e : b200sk_last_error(s->slot[1].ctx);
}

void b200sk_fxstream_close(b200sk_fxstream *s) {
    if (!s) return;
    {
        std::unique_lock<std::mutex> lk(s->mu);
        s->exhausted = true;
        s->cv.wait(lk, [&] { return s->slot[0].state != 1 && s->slot[1].state != 1; });
        s->stop = true;
        s->cv.notify_all();
    }
    for (auto &sl : s->slot) {
        if (sl.th.joinable()) sl.th.join();
        b200sk_destroy(sl.ctx);
    }
    delete s;
}

} // extern "C"
