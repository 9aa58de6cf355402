// Package colour: addition for the GPU worker backend (drop into shared/colour/ of MWindels/distributed-raytracer).
// NOT COMPILED in the repository this file ships in (no Go toolchain there); see integration/go/README.md.
package colour

// Floats exposes the three channels of an RGB object (each in [0, 1]); the fields are unexported (colour.go:16-18).
func (rgb RGB) Floats() (float64, float64, float64) {
	return rgb.r, rgb.g, rgb.b
}
