/* wire_decode.c -- see include/rt_wire.h.  Plain C99, no CUDA, no allocation. */
#include <stddef.h>
#include <string.h>

#include "../../include/rt_wire.h"

/* base-128 varint (up to 10 bytes); returns the new position or 0 on error */
static uint64_t get_varint(const uint8_t *p, uint64_t pos, uint64_t end, uint64_t *out) {
    uint64_t v = 0;
    for (int shift = 0; shift < 70; shift += 7) {
        if (pos >= end) return 0;
        uint8_t b = p[pos++];
        if (shift < 64) v |= (uint64_t)(b & 0x7f) << shift;
        if (!(b & 0x80)) { *out = v; return pos; }
    }
    return 0;
}

/* a field tag is a 32-bit varint with a non-zero field number (field numbers are at most 2^29 - 1) */
static uint64_t get_tag(const uint8_t *p, uint64_t pos, uint64_t end, uint64_t *field, unsigned *wt) {
    uint64_t tag;
    pos = get_varint(p, pos, end, &tag);
    if (!pos || tag > 0xFFFFFFFFull || (tag >> 3) == 0) return 0;
    *field = tag >> 3;
    *wt = (unsigned)(tag & 7);
    return pos;
}

/* skips one field of wire type wt (field number `field`, needed to match the end of a group); returns the new position
 * or 0 on error.  Unknown groups (wire types 3 / 4, proto2 leftovers) are skipped like protobuf runtimes do. */
static uint64_t skip_field(const uint8_t *p, uint64_t pos, uint64_t end, uint64_t field, unsigned wt, int depth) {
    uint64_t v;
    switch (wt) {
    case 0: return get_varint(p, pos, end, &v);
    case 1: return end - pos >= 8 ? pos + 8 : 0;
    case 2:
        pos = get_varint(p, pos, end, &v);
        if (!pos || v > end - pos) return 0;
        return pos + v;
    case 3:
        if (depth > 64) return 0;
        for (;;) {
            uint64_t f; unsigned w;
            pos = get_tag(p, pos, end, &f, &w);
            if (!pos) return 0;
            if (w == 4) return f == field ? pos : 0;
            pos = skip_field(p, pos, end, f, w, depth + 1);
            if (!pos) return 0;
        }
    case 5: return end - pos >= 4 ? pos + 4 : 0;
    default: return 0;               /* a stray end-group, wire types 6 and 7 */
    }
}

/* one Colour sub-message [pos, end): fields 1..3 as varints, last occurrence wins, everything else skipped */
static int parse_colour(const uint8_t *p, uint64_t pos, uint64_t end, uint8_t out[3]) {
    out[0] = out[1] = out[2] = 0;                         /* proto3 default */
    while (pos < end) {
        uint64_t field, v;
        unsigned wt;
        pos = get_tag(p, pos, end, &field, &wt);
        if (!pos) return -1;
        if (field >= 1 && field <= 3 && wt == 0) {
            pos = get_varint(p, pos, end, &v);
            if (!pos) return -1;
            out[field - 1] = (uint8_t)v;                  /* uint8(uint32(v)): master/main.go:172 */
        } else {                                          /* unknown field, or a known one with another wire type */
            pos = skip_field(p, pos, end, field, wt, 0);
            if (!pos) return -1;
        }
    }
    return 0;
}

struct pack_ctx { uint8_t *rgb; uint64_t max; };
struct blit_ctx { uint8_t *surface; uint32_t x, y, w, h, pitch, rs, gs, bs, amask; };

static inline int emit(struct pack_ctx *p, struct blit_ctx *b, uint64_t k, const uint8_t rgb[3]) {
    if (p) {
        if (k >= p->max) return RT_WIRE_ERR_CAPACITY;
        if (p->rgb) { p->rgb[3 * k] = rgb[0]; p->rgb[3 * k + 1] = rgb[1]; p->rgb[3 * k + 2] = rgb[2]; }
        return 0;
    }
    if (k >= (uint64_t)b->w * b->h) return RT_WIRE_ERR_CAPACITY;
    const uint32_t i = (uint32_t)(k / b->h), j = (uint32_t)(k % b->h);             /* results[i*height + j] */
    const uint32_t v = ((uint32_t)rgb[0] << b->rs) | ((uint32_t)rgb[1] << b->gs) | ((uint32_t)rgb[2] << b->bs) | b->amask;
    memcpy(b->surface + (size_t)(b->y + j) * b->pitch + (size_t)(b->x + i) * 4, &v, 4);
    return 0;
}

static int64_t walk(const uint8_t *wire, uint64_t len, struct pack_ctx *pc, struct blit_ctx *bc) {
    uint64_t pos = 0, k = 0;
    while (pos < len) {
        /* fast path: the canonical element every encoder emits -- 0x0A, 1-byte length, fields 1..3 in 1-byte tags with
         * 1- or 2-byte values (90 % of a frame is the 2-byte empty Colour).  Anything else takes the general parser. */
        if (wire[pos] == 0x0A && pos + 1 < len && wire[pos + 1] < 0x80 && wire[pos + 1] <= len - pos - 2) {
            const uint8_t *q = wire + pos + 2, *qe = q + wire[pos + 1];
            uint8_t rgb[3] = { 0, 0, 0 };
            int canonical = 1;
            while (q < qe) {
                const uint8_t tag = *q;
                const unsigned f = tag >> 3;
                if ((tag & 0x87) == 0 && f >= 1 && f <= 3 && q + 1 < qe) {
                    const uint8_t b0 = q[1];
                    if (b0 < 0x80) { rgb[f - 1] = b0; q += 2; continue; }
                    if (q + 2 < qe && q[2] < 0x80) { rgb[f - 1] = (uint8_t)((b0 & 0x7f) | (q[2] << 7)); q += 3; continue; }
                }
                canonical = 0;
                break;
            }
            if (canonical) {
                int rc = emit(pc, bc, k, rgb);
                if (rc != 0) return rc;
                k++;
                pos += 2u + wire[pos + 1];
                continue;
            }
        }
        uint64_t field, n;
        unsigned wt;
        pos = get_tag(wire, pos, len, &field, &wt);
        if (!pos) return RT_WIRE_ERR_MALFORMED;
        if (field == 1 && wt == 2) {
            pos = get_varint(wire, pos, len, &n);
            if (!pos || n > len - pos) return RT_WIRE_ERR_MALFORMED;
            uint8_t rgb[3];
            if (parse_colour(wire, pos, pos + n, rgb) != 0) return RT_WIRE_ERR_MALFORMED;
            int rc = emit(pc, bc, k, rgb);
            if (rc != 0) return rc;
            k++;
            pos += n;
        } else {
            pos = skip_field(wire, pos, len, field, wt, 0);
            if (!pos) return RT_WIRE_ERR_MALFORMED;
        }
    }
    return (int64_t)k;
}

int64_t rt_wire_decode_trace_results(const uint8_t *wire, uint64_t len, uint8_t *rgb, uint64_t max_results) {
    if (!wire && len) return RT_WIRE_ERR_ARGUMENT;
    struct pack_ctx c = { rgb, max_results };
    return walk(wire, len, &c, NULL);
}

int rt_wire_blit_trace_results(const uint8_t *wire, uint64_t len, uint32_t x, uint32_t y, uint32_t w, uint32_t h, void *surface,
                               uint32_t surf_w, uint32_t surf_h, uint32_t pitch, uint32_t rshift, uint32_t gshift, uint32_t bshift,
                               uint32_t amask) {
    if (!surface || (!wire && len) || rshift > 24 || gshift > 24 || bshift > 24) return RT_WIRE_ERR_ARGUMENT;
    if ((uint64_t)x + w > surf_w || (uint64_t)y + h > surf_h || (uint64_t)surf_w * 4 > pitch) return RT_WIRE_ERR_ARGUMENT;
    struct blit_ctx c = { (uint8_t *)surface, x, y, w, h, pitch, rshift, gshift, bshift, amask };
    int64_t n = walk(wire, len, NULL, &c);
    if (n < 0) return (int)n;
    return n == (int64_t)((uint64_t)w * h) ? 0 : RT_WIRE_ERR_MALFORMED;            /* fewer results than the order has pixels */
}
