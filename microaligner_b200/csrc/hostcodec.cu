// Host-side helpers of the TIFF page reader (microaligner_b200/tiffio.py): the TIFF flavour of LZW.
// The reference reads pages through tifffile, whose LZW decoder lives in the compiled `imagecodecs` extension;
// a pure-Python decoder manages ~1 MB/s, this one a few hundred.  No CUDA here -- plain C++ in the same library.
#include <cstdint>
#include <cstring>
#include "../../include/microaligner_b200.h"

// TIFF LZW (TIFF 6.0 section 13): MSB-first codes of 9..12 bits, ClearCode 256, EndOfInformation 257, code width
// grows one code early ("early change").  Strings are stored as (prefix code, last byte, length) and unrolled
// backwards into the output.  Returns the number of bytes written, or -1 on a corrupt stream.
extern "C" long long ma_tiff_lzw_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    if (!src || !dst) return -1;
    static const int kMax = 4096;
    uint16_t prefix[kMax];
    uint8_t last[kMax], first[kMax];
    uint16_t length[kMax];
    for (int i = 0; i < 256; ++i) { prefix[i] = 0; last[i] = first[i] = (uint8_t)i; length[i] = 1; }
    int next = 258, width = 9, prev = -1;
    uint32_t bits = 0;
    int nbits = 0;
    size_t pos = 0, out = 0;
    while (out < cap) {
        while (nbits < width && pos < n) { bits = (bits << 8) | src[pos++]; nbits += 8; }
        if (nbits < width) break;
        const int code = (int)((bits >> (nbits - width)) & ((1u << width) - 1));
        nbits -= width;
        if (code == 257) break;
        if (code == 256) { next = 258; width = 9; prev = -1; continue; }
        int len;
        if (prev < 0) {
            if (code >= 256) return -1;
            len = 1;
            dst[out] = (uint8_t)code;
        } else {
            if (code > next || next >= kMax + 1) return -1;
            const bool known = code < next;
            const int src_code = known ? code : prev;           // KwKwK: the new string is prev + first(prev)
            len = length[src_code] + (known ? 0 : 1);
            // unroll src_code backwards into dst[out .. out + length)
            size_t end = out + length[src_code];
            int c = src_code;
            size_t p = end;
            while (true) {
                --p;
                if (p < cap) dst[p] = last[c];
                if (length[c] == 1) break;
                c = prefix[c];
            }
            if (!known && end < cap) dst[end] = first[prev];
            if (next < kMax) {
                prefix[next] = (uint16_t)prev;
                last[next] = known ? first[code] : first[prev];
                first[next] = first[prev];
                length[next] = (uint16_t)(length[prev] + 1);
                ++next;
            }
        }
        out += (size_t)len;
        prev = code;
        if (next >= (1 << width) - 1 && width < 12) ++width;
    }
    return (long long)(out < cap ? out : cap);
}
