// Package rtb200 binds librt_b200.so (include/rt_b200.h) -- the B200 render backend -- for the Go worker.
// Drop into worker/gpu/rtb200/ of MWindels/distributed-raytracer.  NOT COMPILED in the repository this file ships in
// (no Go toolchain there); the identical C ABI is exercised from C++ and Python ctypes in that repository's tests.
package rtb200

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../lib -lrt_b200 -Wl,-rpath,${SRCDIR}/../../../lib
#include <stdlib.h>
#include "rt_b200.h"
*/
import "C"

import (
	"fmt"
	"math"
	"runtime"
	"unsafe"

	"github.com/mwindels/distributed-raytracer/shared/state"
)

// Scene is one registered scene on one GPU (rt_scene): immutable geometry + BVH + a pool of lanes, so BulkTrace may be
// called from any number of goroutines at once (grpc-go serves every RPC on its own goroutine).
type Scene struct {
	h   *C.rt_scene
	ids []uint // object ids in the order their positions are handed over
}

// NewScene uploads the immutable scene and builds the BVHs on `device` (what register() does once per registration,
// worker/distributed/main.go:101-129).  fast selects RT_FP32_FAST (fp32 traversal, fp64 shading) over RT_FP64_EXACT.
func NewScene(env state.Environment, device int, fast bool) (*Scene, error) {
	flat := env.Flatten()
	n := len(flat.Meshes)
	// rt_mesh_desc holds C pointers into Go memory: allowed for the duration of the call because the slices contain no
	// Go pointers; the descriptor array itself lives in C memory, and the slices are pinned while C reads them.
	var pin runtime.Pinner
	defer pin.Unpin()
	descs := (*C.rt_mesh_desc)(C.calloc(C.size_t(maxInt(n, 1)), C.sizeof_rt_mesh_desc))
	defer C.free(unsafe.Pointer(descs))
	dv := unsafe.Slice(descs, maxInt(n, 1))
	for i := range flat.Meshes {
		f := &flat.Meshes[i]
		if len(f.Verts) > 0 {
			pin.Pin(&f.Verts[0])
			dv[i].verts = (*C.double)(&f.Verts[0])
		}
		dv[i].n_verts = C.uint32_t(len(f.Verts) / 3)
		if len(f.Normals) > 0 {
			pin.Pin(&f.Normals[0])
			dv[i].normals = (*C.double)(&f.Normals[0])
		}
		dv[i].n_normals = C.uint32_t(len(f.Normals) / 3)
		if len(f.Faces) > 0 {
			pin.Pin(&f.Faces[0])
			dv[i].faces = (*C.uint32_t)(&f.Faces[0])
		}
		dv[i].n_faces = C.uint32_t(len(f.Faces) / 7)
		if len(f.Materials) > 0 {
			pin.Pin(&f.Materials[0])
			dv[i].materials = (*C.double)(&f.Materials[0])
		}
		dv[i].n_materials = C.uint32_t(len(f.Materials) / 10)
	}
	var d C.rt_scene_desc
	d.meshes, d.n_meshes = descs, C.uint32_t(n)
	if len(flat.ObjMesh) > 0 {
		pin.Pin(&flat.ObjPos[0])
		pin.Pin(&flat.ObjMesh[0])
		d.obj_pos, d.obj_mesh = (*C.double)(&flat.ObjPos[0]), (*C.uint32_t)(&flat.ObjMesh[0])
	}
	d.n_objects = C.uint32_t(len(flat.ObjMesh))
	d.device = C.int32_t(device)
	d.precision = C.RT_FP64_EXACT
	if fast {
		d.precision = C.RT_FP32_FAST
	}
	s := &Scene{ids: flat.ObjIds}
	if rc := C.rt_scene_create(&d, &s.h); rc != 0 {
		return nil, fmt.Errorf("rt_scene_create: %d %s", int(rc), C.GoString(C.rt_last_error(nil)))
	}
	runtime.SetFinalizer(s, func(s *Scene) { s.Close() })
	return s, nil
}

// Close releases the GPU scene (re-registration: Close the old scene, NewScene the new one; main.go:146-185).
func (s *Scene) Close() {
	if s.h != nil {
		C.rt_scene_destroy(s.h)
		s.h = nil
	}
}

type frame struct {
	cam    C.rt_camera
	lights []C.rt_light
	pos    []float64
}

func (s *Scene) frameOf(em *state.EnvMutables) frame {
	ff := em.Flatten(s.ids)
	var f frame
	for k := 0; k < 3; k++ {
		f.cam.pos[k], f.cam.forward[k] = C.double(ff.Cam[k]), C.double(ff.Cam[3+k])
		f.cam.left[k], f.cam.up[k] = C.double(ff.Cam[6+k]), C.double(ff.Cam[9+k])
	}
	f.cam.fov = C.double(ff.Cam[12])
	// Go's own math.Tan: the frame constant of tracer.go:17, bit for bit what the reference computes per pixel
	f.cam.tan_half_fov = C.double(math.Tan(ff.Cam[12] / 2.0))
	f.lights = make([]C.rt_light, len(ff.Lights)/6)
	for i := range f.lights {
		for k := 0; k < 3; k++ {
			f.lights[i].pos[k], f.lights[i].col[k] = C.double(ff.Lights[6*i+k]), C.double(ff.Lights[6*i+3+k])
		}
	}
	f.pos = ff.ObjPos
	return f
}

func (f *frame) ptrs() (*C.rt_light, *C.double) {
	var lp *C.rt_light
	var pp *C.double
	if len(f.lights) > 0 {
		lp = &f.lights[0]
	}
	if len(f.pos) > 0 {
		pp = (*C.double)(&f.pos[0])
	}
	return lp, pp
}

// BulkTrace renders region (x, y, w, h) of a screenW x screenH screen for the frame state em (already LinkTo'd).
// out must hold w*h*3 bytes; pixel (x+i, y+j) lands at out[(i*h+j)*3:] -- the index BulkTrace uses (main.go:82) --
// quantised like colour.RGB() (colour.go:59-61).  Blocks until the bytes are in out; safe to call concurrently.
func (s *Scene) BulkTrace(em *state.EnvMutables, screenW, screenH, x, y, w, h uint32, out []byte) error {
	if uint64(len(out)) < uint64(w)*uint64(h)*3 {
		return fmt.Errorf("rtb200: out holds %d bytes, %d needed", len(out), uint64(w)*uint64(h)*3)
	}
	if w == 0 || h == 0 {
		return nil
	}
	f := s.frameOf(em)
	lp, pp := f.ptrs()
	rc := C.rt_bulk_trace(s.h, &f.cam, lp, C.uint32_t(len(f.lights)), pp,
		C.uint32_t(screenW), C.uint32_t(screenH), C.uint32_t(x), C.uint32_t(y), C.uint32_t(w), C.uint32_t(h),
		(*C.uint8_t)(&out[0]), C.RT_LAYOUT_COLMAJOR|C.RT_QUANT_RGB255)
	runtime.KeepAlive(f)
	if rc != 0 {
		return fmt.Errorf("rt_bulk_trace: %d %s", int(rc), C.GoString(C.rt_last_error(s.h)))
	}
	return nil
}

// BulkTraceEncoded answers the same order with the protobuf wire bytes of comms.TraceResults (comms.proto:34-41):
// exactly what proto.Marshal would emit for the message BulkTrace builds (main.go:52-54,79-86), encoded on the GPU.
func (s *Scene) BulkTraceEncoded(em *state.EnvMutables, screenW, screenH, x, y, w, h uint32) ([]byte, error) {
	if w == 0 || h == 0 {
		return nil, nil // the empty message
	}
	f := s.frameOf(em)
	lp, pp := f.ptrs()
	buf := make([]byte, uint64(C.rt_trace_results_max_bytes(C.uint64_t(uint64(w)*uint64(h)))))
	var n C.uint64_t
	rc := C.rt_bulk_trace_encoded(s.h, &f.cam, lp, C.uint32_t(len(f.lights)), pp,
		C.uint32_t(screenW), C.uint32_t(screenH), C.uint32_t(x), C.uint32_t(y), C.uint32_t(w), C.uint32_t(h),
		(*C.uint8_t)(&buf[0]), C.uint64_t(len(buf)), &n, C.RT_QUANT_RGB255)
	runtime.KeepAlive(f)
	if rc != 0 {
		return nil, fmt.Errorf("rt_bulk_trace_encoded: %d %s", int(rc), C.GoString(C.rt_last_error(s.h)))
	}
	return buf[:n], nil
}

// Cancel makes renders in flight on this scene return RT_ERR_CANCELLED at their next kernel boundary (ctx.Err(), main.go:73).
func (s *Scene) Cancel() { C.rt_cancel(s.h) }

func maxInt(a, b int) int {
	if a > b {
		return a
	}
	return b
}
