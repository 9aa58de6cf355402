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
	var d C.rt_scene_desc
	release, ids, err := fillSceneDesc(env, fast, &d)
	if err != nil {
		return nil, err
	}
	defer release()
	d.device = C.int32_t(device)
	s := &Scene{ids: ids}
	if rc := C.rt_scene_create(&d, &s.h); rc != 0 {
		return nil, fmt.Errorf("rt_scene_create: %d %s", int(rc), C.GoString(C.rt_last_error(nil)))
	}
	runtime.SetFinalizer(s, func(s *Scene) { s.Close() })
	return s, nil
}

// ObjModel is one model handed over as the TEXT of its OBJ file (rt_obj_mesh_desc): the GPU parses the geometry (v / vn / f lines,
// gwob's semantics, mesh.go:109-133,166-204); the worker only supplies what the small MTL file says -- the material table in
// first-use order (mesh.go:151-163) and, for every `usemtl` line of the text, its byte offset and the material it selects.
type ObjModel struct {
	Text            []byte
	Materials       []float64 // 10 per material: Ka Kd Ks Ns
	UsemtlOffset    []uint64  // ascending byte offsets of the `usemtl` lines
	UsemtlMaterial  []uint32
	DefaultMaterial uint32 // material of the faces above the first usemtl line
}

// NewSceneFromOBJ is NewScene for a worker that has the model FILES (or their bytes from the master) instead of decoded meshes:
// nothing but the files' bytes crosses PCIe, and neither gwob's tokenizer nor MeshFromFile's value-keyed de-duplication maps
// (mesh.go:146-203) nor the per-face R-tree Insert (mesh.go:208) runs on the host.  objPos: 3 per object; objModel: index into models.
func NewSceneFromOBJ(models []ObjModel, objPos []float64, objModel []uint32, ids []uint, device int, fast bool) (*Scene, error) {
	if len(objPos) != 3*len(objModel) {
		return nil, fmt.Errorf("rtb200: %d positions for %d objects", len(objPos)/3, len(objModel))
	}
	pin := new(runtime.Pinner)
	defer pin.Unpin()
	descs := (*C.rt_obj_mesh_desc)(C.calloc(C.size_t(maxInt(len(models), 1)), C.sizeof_rt_obj_mesh_desc))
	defer C.free(unsafe.Pointer(descs))
	dv := unsafe.Slice(descs, maxInt(len(models), 1))
	for i := range models {
		m := &models[i]
		if len(m.Text) > 0 {
			pin.Pin(&m.Text[0])
			dv[i].obj_text = (*C.char)(unsafe.Pointer(&m.Text[0]))
		}
		dv[i].obj_bytes = C.uint64_t(len(m.Text))
		if len(m.Materials) > 0 {
			pin.Pin(&m.Materials[0])
			dv[i].materials = (*C.double)(&m.Materials[0])
		}
		dv[i].n_materials = C.uint32_t(len(m.Materials) / 10)
		if len(m.UsemtlOffset) > 0 {
			pin.Pin(&m.UsemtlOffset[0])
			pin.Pin(&m.UsemtlMaterial[0])
			dv[i].usemtl_offset = (*C.uint64_t)(&m.UsemtlOffset[0])
			dv[i].usemtl_material = (*C.uint32_t)(&m.UsemtlMaterial[0])
		}
		dv[i].n_usemtl = C.uint32_t(len(m.UsemtlOffset))
		dv[i].default_material = C.uint32_t(m.DefaultMaterial)
	}
	var pp *C.double
	var pm *C.uint32_t
	if len(objModel) > 0 {
		pin.Pin(&objPos[0])
		pin.Pin(&objModel[0])
		pp, pm = (*C.double)(&objPos[0]), (*C.uint32_t)(&objModel[0])
	}
	precision := C.uint32_t(C.RT_FP64_EXACT)
	if fast {
		precision = C.uint32_t(C.RT_FP32_FAST)
	}
	s := &Scene{ids: ids}
	if rc := C.rt_scene_create_obj(descs, C.uint32_t(len(models)), pp, pm, C.uint32_t(len(objModel)), C.int32_t(device), precision, &s.h); rc != 0 {
		return nil, fmt.Errorf("rt_scene_create_obj: %d %s", int(rc), C.GoString(C.rt_last_error(nil)))
	}
	runtime.SetFinalizer(s, func(s *Scene) { s.Close() })
	return s, nil
}

// fillSceneDesc flattens env into an rt_scene_desc.  rt_mesh_desc holds C pointers into Go memory: allowed for the duration
// of the create call because the slices contain no Go pointers; the descriptor array itself lives in C memory, and the
// slices stay pinned until release() is called (after rt_scene_create / rt_group_create has copied everything).
func fillSceneDesc(env state.Environment, fast bool, d *C.rt_scene_desc) (release func(), ids []uint, err error) {
	flat := env.Flatten()
	n := len(flat.Meshes)
	pin := new(runtime.Pinner)
	descs := (*C.rt_mesh_desc)(C.calloc(C.size_t(maxInt(n, 1)), C.sizeof_rt_mesh_desc))
	release = func() { pin.Unpin(); C.free(unsafe.Pointer(descs)); runtime.KeepAlive(flat) }
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
	d.meshes, d.n_meshes = descs, C.uint32_t(n)
	if len(flat.ObjMesh) > 0 {
		pin.Pin(&flat.ObjPos[0])
		pin.Pin(&flat.ObjMesh[0])
		d.obj_pos, d.obj_mesh = (*C.double)(&flat.ObjPos[0]), (*C.uint32_t)(&flat.ObjMesh[0])
	}
	d.n_objects = C.uint32_t(len(flat.ObjMesh))
	d.precision = C.RT_FP64_EXACT
	if fast {
		d.precision = C.RT_FP32_FAST
	}
	return release, flat.ObjIds, nil
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

// Cancel stops the renders that are in flight on this scene (ctx.Err(), main.go:73): their kernels poll a cancel generation
// and give up, the blocked BulkTrace calls return RT_ERR_CANCELLED.  Calls that start afterwards are not affected.
func (s *Scene) Cancel() { C.rt_cancel(s.h) }

// Group is one registered scene on SEVERAL GPUs of the box (rt_group): one worker process, one BulkTrace per order, all GPUs
// rendering stripes of it and copying them into the reply over their own PCIe links.  master/pool still sees one worker.
type Group struct {
	h   *C.rt_group
	ids []uint
}

// NewGroup registers the scene on the listed CUDA devices (nil: every GPU of the box).
func NewGroup(env state.Environment, devices []int32, fast bool) (*Group, error) {
	var d C.rt_scene_desc
	release, ids, err := fillSceneDesc(env, fast, &d) // the body of NewScene up to rt_scene_create, shared by both constructors
	if err != nil {
		return nil, err
	}
	defer release()
	g := &Group{ids: ids}
	var dp *C.int32_t
	if len(devices) > 0 {
		dp = (*C.int32_t)(&devices[0])
	}
	if rc := C.rt_group_create(&d, dp, C.uint32_t(len(devices)), &g.h); rc != 0 {
		return nil, fmt.Errorf("rt_group_create: %d %s", int(rc), C.GoString(C.rt_group_last_error(nil)))
	}
	runtime.SetFinalizer(g, func(g *Group) { g.Close() })
	return g, nil
}

func (g *Group) Close() {
	if g.h != nil {
		C.rt_group_destroy(g.h)
		g.h = nil
	}
}

// BulkTrace: same contract as Scene.BulkTrace (bytes identical to the single-GPU frame); safe to call concurrently.
// Hand it a buffer from rt_host_alloc (pinned) and the GPUs copy their stripes straight into it.
func (g *Group) BulkTrace(em *state.EnvMutables, screenW, screenH, x, y, w, h uint32, out []byte) error {
	if uint64(len(out)) < uint64(w)*uint64(h)*3 {
		return fmt.Errorf("rtb200: out holds %d bytes, %d needed", len(out), uint64(w)*uint64(h)*3)
	}
	if w == 0 || h == 0 {
		return nil
	}
	f := (&Scene{ids: g.ids}).frameOf(em)
	lp, pp := f.ptrs()
	rc := C.rt_group_bulk_trace(g.h, &f.cam, lp, C.uint32_t(len(f.lights)), pp,
		C.uint32_t(screenW), C.uint32_t(screenH), C.uint32_t(x), C.uint32_t(y), C.uint32_t(w), C.uint32_t(h),
		(*C.uint8_t)(&out[0]), C.RT_LAYOUT_COLMAJOR|C.RT_QUANT_RGB255)
	runtime.KeepAlive(f)
	if rc != 0 {
		return fmt.Errorf("rt_group_bulk_trace: %d %s", int(rc), C.GoString(C.rt_group_last_error(g.h)))
	}
	return nil
}

func maxInt(a, b int) int {
	if a > b {
		return a
	}
	return b
}
