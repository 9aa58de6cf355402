// Package state: the one permitted touch of shared/state for the GPU worker backend -- an exported flattener that
// turns the (mostly unexported) scene data into plain slices the C ABI of include/rt_b200.h can take.
// Drop into shared/state/ of MWindels/distributed-raytracer.  NOT COMPILED in the repository this file ships in.
package state

import (
	"math"
	"sort"

	"github.com/mwindels/rtreego"
)

// FlatMesh is the structure-of-arrays image of one Mesh (rt_mesh_desc in include/rt_b200.h).
type FlatMesh struct {
	Verts     []float64 // 3 per vertex            (mesh.go:101)
	Normals   []float64 // 3 per vertex normal, may be empty => flat normals (object.go:82,93-95)
	Faces     []uint32  // 7 per face: v0 v1 v2 n0 n1 n2 mat (mesh.go:21-27)
	Materials []float64 // 10 per material: Ka Kd Ks Ns      (mesh.go:94-97)
}

// Flatten lists the faces in R-tree enumeration order -- the only order the reference can observe (mesh.go:228) --
// which makes "first candidate wins on a distance tie" (object.go:97) the same as "lowest face index wins".
func (m *Mesh) Flatten() FlatMesh {
	var f FlatMesh
	for _, v := range m.vertices {
		f.Verts = append(f.Verts, v.X, v.Y, v.Z)
	}
	for _, n := range m.vertexNormals {
		f.Normals = append(f.Normals, n.X, n.Y, n.Z)
	}
	for _, mt := range m.materials {
		ar, ag, ab := mt.Ka.Floats()
		dr, dg, db := mt.Kd.Floats()
		sr, sg, sb := mt.Ks.Floats()
		f.Materials = append(f.Materials, ar, ag, ab, dr, dg, db, sr, sg, sb, mt.Ns)
	}
	for _, s := range m.faces.SearchCondition(func(*rtreego.Rect) bool { return true }) {
		fc := s.(face)
		f.Faces = append(f.Faces,
			uint32(fc.verts[0]), uint32(fc.verts[1]), uint32(fc.verts[2]),
			uint32(fc.vertNorms[0]), uint32(fc.vertNorms[1]), uint32(fc.vertNorms[2]), uint32(fc.mat))
	}
	return f
}

// FlatScene is the immutable part of an environment as the GPU backend registers it (rt_scene_desc):
// the distinct meshes, and per object (in ascending id order) its mesh index.
type FlatScene struct {
	Meshes  []FlatMesh
	ObjIds  []uint   // ascending; FlatFrame lists positions in the same order
	ObjMesh []uint32 // index into Meshes
	ObjPos  []float64
}

// Flatten walks the immutable part (environment.go:25-28) and the current objects.
func (e Environment) Flatten() FlatScene {
	var fs FlatScene
	paths := make([]string, 0, len(e.immutable.meshes))
	for p := range e.immutable.meshes {
		paths = append(paths, p)
	}
	sort.Strings(paths) // map order is random; the mesh indices must be reproducible
	index := make(map[string]uint32, len(paths))
	for i, p := range paths {
		index[p] = uint32(i)
		fs.Meshes = append(fs.Meshes, e.immutable.meshes[p].Flatten())
	}
	for _, o := range e.mutable.sortedObjects() {
		path, ok := e.immutable.paths[o.id]
		if !ok {
			continue // an object without a model has no mesh (environment.go:84-90): it can never be hit
		}
		fs.ObjIds = append(fs.ObjIds, o.id)
		fs.ObjMesh = append(fs.ObjMesh, index[path])
		fs.ObjPos = append(fs.ObjPos, o.Pos.X, o.Pos.Y, o.Pos.Z)
	}
	return fs
}

// FlatFrame is the per-order mutable state (the gob `diff` of a WorkOrder, distributed/main.go:56-64).
type FlatFrame struct {
	Cam    [13]float64 // pos, forward, left, up (3 each), fov
	Lights []float64   // 6 per light: pos, colour
	ObjPos []float64   // 3 per object, in the id order of FlatScene.ObjIds
}

// Flatten lists the frame state; ids must be the FlatScene.ObjIds of the registered scene.
func (em *EnvMutables) Flatten(ids []uint) FlatFrame {
	var ff FlatFrame
	c := em.Cam
	copy(ff.Cam[0:3], []float64{c.Pos.X, c.Pos.Y, c.Pos.Z})
	copy(ff.Cam[3:6], []float64{c.forward.X, c.forward.Y, c.forward.Z})
	copy(ff.Cam[6:9], []float64{c.left.X, c.left.Y, c.left.Z})
	copy(ff.Cam[9:12], []float64{c.up.X, c.up.Y, c.up.Z})
	ff.Cam[12] = c.Fov
	for _, l := range em.Lights {
		r, g, b := l.Col.Floats()
		ff.Lights = append(ff.Lights, l.Pos.X, l.Pos.Y, l.Pos.Z, r, g, b)
	}
	pos := make(map[uint]*Object)
	for _, o := range em.sortedObjects() {
		pos[o.id] = o
	}
	for _, id := range ids {
		if o, ok := pos[id]; ok {
			ff.ObjPos = append(ff.ObjPos, o.Pos.X, o.Pos.Y, o.Pos.Z)
		} else {
			// not in this order's diff: LinkTo (environment.go:73-98) rebuilds the object tree from the diff alone, so the
			// reference does not render it.  A NaN position marks the object absent for the backend (rt_b200.h).
			ff.ObjPos = append(ff.ObjPos, math.NaN(), math.NaN(), math.NaN())
		}
	}
	return ff
}

// sortedObjects enumerates the object R-tree (environment.go:76) in ascending id order.
func (em *EnvMutables) sortedObjects() []*Object {
	var out []*Object
	if em.Objs == nil {
		return out
	}
	for _, s := range em.Objs.SearchCondition(func(*rtreego.Rect) bool { return true }) {
		out = append(out, s.(*Object))
	}
	sort.Slice(out, func(i, j int) bool { return out[i].id < out[j].id })
	return out
}
