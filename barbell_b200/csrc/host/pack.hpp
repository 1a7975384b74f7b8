// Host-side nibble packing of read bases for the host->device copy (see pack.cpp).
#pragma once
#include <cstddef>
#include <cstdint>

namespace bb {
// code[256]: byte -> 4-bit IUPAC base set (the only property of a text byte the search depends on).
// Packs n bases into (n+1)/2 bytes: out[i] = code[src[2i]] | code[src[2i+1]] << 4, using up to `threads` host threads.
// returns the seconds spent packing (the wait for the shared thread pool, when another caller is packing, is not counted)
double pack_nibbles(const uint8_t* src, size_t n, uint8_t* dst, const uint8_t* code, int threads);
// Two bits per base for A/C/G/T (any case, U as T) into (n+3)/4 bytes, every other byte as an exception entry
// (position << 4 | base set) appended to exc[0, exc_cap) in blocks; unused entries of a block hold ~0.  *n_exc = entries to copy,
// *overflow = the list was too small (the output is then unusable).  Returns the packing seconds like pack_nibbles.
double pack_crumbs(const uint8_t* src, size_t n, uint8_t* dst, uint64_t* exc, size_t exc_cap, size_t* n_exc, bool* overflow,
                   const uint8_t* code, int threads);
// Appends n bases to a gapless 2-bit stream (base i lives in dst[i >> 2], bits 2 * (i & 3)): *pos = bases in the stream so far, updated;
// exceptions are appended at exc[*n_exc ...) with their position IN THE STREAM.  One writer per stream; false = exception list full.
// (For producers that pack while they parse: the bytes of a read are touched once, while they are still in the cache.)
bool crumbs_append(const uint8_t* src, size_t n, uint8_t* dst, uint64_t* pos, uint64_t* exc, size_t exc_cap, size_t* n_exc, const uint8_t* code);
// The same for a text line of unknown length: packs src[0, e) where e = index of the first '\n' in src[0, n) (*found) or n; a '\r' in front
// of the '\n' is not part of the line.  *line_len = e.  One pass over the bytes instead of memchr + pack.
bool crumbs_append_line(const uint8_t* src, size_t n, uint8_t* dst, uint64_t* pos, uint64_t* exc, size_t exc_cap, size_t* n_exc, const uint8_t* code,
                        size_t* line_len, bool* found);
int pack_default_threads();
}  // namespace bb
