// partition_balanced.go -- SURVEY s8 row f4.  Drop into master/ of MWindels/distributed-raytracer next to main.go.
// NOT COMPILED in the repository this file ships in (no Go toolchain there); the identical algorithm is implemented and
// tested in C++ (distributed-raytracer_b200/host/gpu_worker.cpp: worker::partitionBalanced, tests/test_host.py).
//
// partition() (master/main.go:54-91) halves the area; with GPU workers a frame costs what its HIT pixels cost (a hit pixel of the
// example scene is ~30x a miss: one nearest-hit traversal plus a shadow ray per light), so equal areas are unequal work -- on the
// 120-frame orbit of config C5 eight equal rectangles leave most workers idle behind the two that hold the model.  This variant
// keeps partition()'s recursion, worker accounting and kernel limits, and only moves every CUT to where the cost of the
// PREVIOUS frame splits in proportion to the workers on either side.  The previous frame is what newCoordinator already holds
// when it draws (orderMap, master/main.go:166-176); no protocol change, the workers are untouched.
package main

import (
	"github.com/mwindels/distributed-raytracer/shared/comms"
)

// hitWeight is the cost of a pixel that hit something, in units of the cost of a pixel that did not.
const hitWeight = 30.0

// costTable is the summed-area table of last frame's per-pixel cost, column-major like a TraceResults reply:
// sat[i*(h+1)+j] = cost of the pixels [0,i) x [0,j).
type costTable struct {
	w, h uint32
	sat  []float64
}

// newCostTable builds the table from the replies the coordinator has just drawn: a pixel costs 1, a lit one 1 + hitWeight.
func newCostTable(width, height uint32, orderMap map[*comms.WorkOrder]*comms.TraceResults) *costTable {
	ct := &costTable{w: width, h: height, sat: make([]float64, (uint64(width)+1)*(uint64(height)+1))}
	hit := make([]bool, uint64(width)*uint64(height))
	for o, r := range orderMap {
		pixels := r.GetResults()
		x0, y0, w, h := o.GetX(), o.GetY(), o.GetWidth(), o.GetHeight()
		for i := uint32(0); i < w; i++ {
			for j := uint32(0); j < h; j++ {
				p := pixels[i*h+j]
				if p.GetR()|p.GetG()|p.GetB() != 0 {
					hit[uint64(x0+i)*uint64(height)+uint64(y0+j)] = true
				}
			}
		}
	}
	stride := uint64(height) + 1
	for i := uint32(0); i < width; i++ {
		col := 0.0 // running cost of pixel column i
		for j := uint32(0); j < height; j++ {
			col += 1.0
			if hit[uint64(i)*uint64(height)+uint64(j)] {
				col += hitWeight
			}
			ct.sat[uint64(i+1)*stride+uint64(j+1)] = ct.sat[uint64(i)*stride+uint64(j+1)] + col
		}
	}
	return ct
}

func (ct *costTable) sum(x, y, w, h uint32) float64 {
	stride := uint64(ct.h) + 1
	at := func(i, j uint32) float64 { return ct.sat[uint64(i)*stride+uint64(j)] }
	return at(x+w, y+h) - at(x, y+h) - at(x+w, y) + at(x, y)
}

// partitionBalanced is partition() with the cuts placed by cost.  With ct == nil (first frame) it IS partition().
func partitionBalanced(ct *costTable, area *comms.WorkOrder, workers uint, dimension uint) ([]comms.WorkOrder, uint) {
	if ct == nil {
		return partition(area, workers, dimension)
	}
	// master/main.go:56-62
	if workers/workerRedundancy < 2 {
		if workers > workerRedundancy {
			return []comms.WorkOrder{*area}, workers % workerRedundancy
		}
		return []comms.WorkOrder{*area}, 0
	}
	x, y := area.GetX(), area.GetY()
	width, height := area.GetWidth(), area.GetHeight()
	if width <= widthKernel && height <= heightKernel { // :66-68
		return []comms.WorkOrder{*area}, workers - workerRedundancy
	} else if width <= widthKernel {
		dimension = 1
	} else if height <= heightKernel {
		dimension = 0
	}
	workersLeft := workers/2 + workers%2 // :88
	want := ct.sum(x, y, width, height) * float64(workersLeft) / float64(workers)
	vertical := dimension%2 == 0
	extent := height
	if vertical {
		extent = width
	}
	// smallest cut whose left part carries at least its share of the cost (the prefix cost is monotone), at least one pixel
	// away from both ends
	lo, hi := uint32(1), extent-1
	for lo < hi {
		mid := lo + (hi-lo)/2
		var c float64
		if vertical {
			c = ct.sum(x, y, mid, height)
		} else {
			c = ct.sum(x, y, width, mid)
		}
		if c < want {
			lo = mid + 1
		} else {
			hi = mid
		}
	}
	cut := lo
	var leftOrder, rightOrder *comms.WorkOrder
	if vertical {
		leftOrder = &comms.WorkOrder{X: x, Y: y, Width: cut, Height: height, Diff: area.GetDiff()}
		rightOrder = &comms.WorkOrder{X: x + cut, Y: y, Width: width - cut, Height: height, Diff: area.GetDiff()}
	} else {
		leftOrder = &comms.WorkOrder{X: x, Y: y, Width: width, Height: cut, Diff: area.GetDiff()}
		rightOrder = &comms.WorkOrder{X: x, Y: y + cut, Width: width, Height: height - cut, Diff: area.GetDiff()}
	}
	left, remainder := partitionBalanced(ct, leftOrder, workersLeft, (dimension+1)%2)
	right, remainder := partitionBalanced(ct, rightOrder, workers/2+remainder, (dimension+1)%2)
	return append(left, right...), remainder
}

// The one edit to newCoordinator (master/main.go:101-103): keep the table of the frame just drawn in the system struct and use it for
// the next frame's orders,
//
//	orders, _ := partitionBalanced(sys.lastCost, &comms.WorkOrder{X: 0, Y: 0, Width: screenWidth, Height: screenHeight, Diff: diff}, workers, 0)
//	...
//	sys.lastCost = newCostTable(screenWidth, screenHeight, orderMap)   // after the draw loop, :166-176
