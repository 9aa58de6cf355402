/*
 * rt_wire.h -- master-side companion of the GPU worker (second half).  Plain C, NO CUDA.
 *
 * The reference master receives comms.TraceResults (shared/comms/comms.proto:34-41: repeated Colour{uint32 r=1,g=2,b=3}),
 * lets protobuf-go materialise one *TraceResults_Colour per pixel and then calls surface.Set per pixel
 * (master/main.go:166-176).  These functions parse the wire bytes directly -- any valid protobuf encoding of the message,
 * not only the canonical one the GPU encoder emits: fields in any order, repeated scalars (last one wins), unknown
 * fields skipped, non-minimal varints, values above 255 truncated like the master's uint8(pixel.GetR()) -- and write
 * packed pixels.  Built as librt_wire.so (no dependencies) and into librt_host.so.
 */
#ifndef RT_WIRE_H
#define RT_WIRE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define RT_WIRE_ERR_MALFORMED (-1)   /* truncated input, bad varint, group wire type, sub-message overrun */
#define RT_WIRE_ERR_CAPACITY  (-2)   /* more results than the caller allowed for */
#define RT_WIRE_ERR_ARGUMENT  (-3)

/* Results in message order, 3 bytes each (r, g, b); returns the number of results or a negative error.
 * rgb may be NULL to count only. */
int64_t rt_wire_decode_trace_results(const uint8_t *wire, uint64_t len, uint8_t *rgb, uint64_t max_results);

/* The frame-assembly loop of master/main.go:166-176 for one order: result k = i*h + j is pixel (x+i, y+j); it is stored
 * as the 32-bit value (r<<rshift | g<<gshift | b<<bshift | amask) at surface + (y+j)*pitch + (x+i)*4 -- a 32-bit SDL
 * window surface (e.g. XRGB8888: rshift 16, gshift 8, bshift 0).  The message must hold exactly w*h results.
 * Returns 0 or a negative error; on error the surface may be partly written. */
int rt_wire_blit_trace_results(const uint8_t *wire, uint64_t len, uint32_t x, uint32_t y, uint32_t w, uint32_t h, void *surface,
                               uint32_t surf_w, uint32_t surf_h, uint32_t pitch, uint32_t rshift, uint32_t gshift, uint32_t bshift,
                               uint32_t amask);
#ifdef __cplusplus
}
#endif
#endif
