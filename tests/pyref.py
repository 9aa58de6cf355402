"""A second, independent restatement of the reference's per-pixel path -- plain scalar Python written from the Go
sources (NOT from oracle/rt_oracle.c), used only to cross-check the C oracle (tests/test_oracle.py).

CPython floats are IEEE-754 binary64, every + - * / is rounded once (no fused multiply-add, no extended precision) and
math.sqrt is correctly rounded: the same arithmetic Go performs on amd64.  Each function names the Go lines it follows
(paths under /root/reference/).  Candidate search is brute force over the faces in index order: a superset of what the
R-tree search returns, with "first candidate wins on a distance tie" == lowest face index.
"""
import math


# ---- shared/geom/vector.go:14-57 -----------------------------------------------------------------
def add(a, b): return (a[0] + b[0], a[1] + b[1], a[2] + b[2])
def sub(a, b): return (a[0] - b[0], a[1] - b[1], a[2] - b[2])
def scale(a, s): return (s * a[0], s * a[1], s * a[2])
def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
def cross(a, b): return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])
def length(a): return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def norm(a):
    mag = math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])
    if mag == 0.0:                                   # Go: x/0 = +-Inf, 0/0 = NaN (no exception)
        return tuple(math.nan if c == 0.0 else math.copysign(math.inf, c) for c in a)
    return (a[0] / mag, a[1] / mag, a[2] / mag)


# ---- Go stdlib ------------------------------------------------------------------------------------
def go_max(x, y):                                    # math/dim.go Max
    if math.isinf(x) and x > 0 or math.isinf(y) and y > 0: return math.inf
    if math.isnan(x) or math.isnan(y): return math.nan
    if x == 0 and x == y: return y if math.copysign(1.0, x) < 0 else x
    return x if x > y else y


def go_min(x, y):                                    # math/dim.go Min
    if math.isinf(x) and x < 0 or math.isinf(y) and y < 0: return -math.inf
    if math.isnan(x) or math.isnan(y): return math.nan
    if x == 0 and x == y: return x if math.copysign(1.0, x) < 0 else y
    return x if x < y else y


def go_pow_int(x, y):
    """math/pow.go for the cases the render path meets: x >= 0 finite, y > 0 finite with no fractional part
    (Frexp, square-and-multiply on the mantissa with the exponent carried separately, Ldexp)."""
    assert y > 0 and y == math.floor(y) and x >= 0 and math.isfinite(x)
    if x == 1: return 1.0
    if y == 1: return x
    if x == 0: return 0.0
    a1, ae = 1.0, 0
    x1, xe = math.frexp(x)
    i = int(y)
    while i != 0:
        if xe < -(1 << 12) or (1 << 12) < xe:
            ae += xe
            break
        if i & 1:
            a1 *= x1
            ae += xe
        x1 *= x1
        xe <<= 1
        if x1 < 0.5:
            x1 += x1
            xe -= 1
        i >>= 1
    try:
        return math.ldexp(a1, ae)
    except OverflowError:
        return math.inf


# ---- shared/geom/triangle.go:37-77 ----------------------------------------------------------------
def tri_intersection(P1, P2, P3, o, d):
    p1p2, p1p3, nd = sub(P2, P1), sub(P3, P1), scale(d, -1)
    incidence = dot(p1p2, cross(p1p3, nd))
    if incidence != 0.0:
        p1or = sub(o, P1)
        r2 = dot(p1or, cross(p1p3, nd)) / incidence
        if 0.0 <= r2 <= 1.0:
            r3 = dot(p1p2, cross(p1or, nd)) / incidence
            if 0.0 <= r2 + r3 <= 1.0:
                r1 = 1.0 - r2 - r3
                if r1 >= 0.0 and r2 >= 0.0 and r3 >= 0.0:
                    s = dot(p1p2, cross(p1p3, p1or)) / incidence
                    if s >= 0.0:
                        return add(o, scale(d, s)), (r1, r2, r3)
    return None


class Scene:
    def __init__(self, sc):
        self.meshes = []
        for m in sc["meshes"]:
            self.meshes.append({
                "v": [tuple(float(c) for c in r) for r in m["verts"]],
                "n": [tuple(float(c) for c in r) for r in m["normals"]],
                "f": [tuple(int(c) for c in r) for r in m["faces"]],
                "m": [tuple(float(c) for c in r) for r in m["materials"]],
            })
        self.objs = [(tuple(float(c) for c in p), int(k)) for p, k in zip(sc["obj_pos"], sc["obj_mesh"])]
        cam = [float(c) for c in sc["camera"]]
        self.cam_pos, self.fwd, self.left, self.up, self.fov = tuple(cam[0:3]), tuple(cam[3:6]), tuple(cam[6:9]), tuple(cam[9:12]), cam[12]
        self.lights = [(tuple(float(c) for c in l[0:3]), tuple(float(c) for c in l[3:6])) for l in sc["lights"]]

    # shared/state/object.go:63-110
    def object_intersection(self, pos, mesh, o, d):
        o = sub(o, pos)
        best = None
        for fi, f in enumerate(mesh["f"]):
            P1, P2, P3 = mesh["v"][f[0]], mesh["v"][f[1]], mesh["v"][f[2]]
            hit = tri_intersection(P1, P2, P3, o, d)
            if hit is None:
                continue
            ip, (r1, r2, r3) = hit
            if mesh["n"]:
                N1, N2, N3 = mesh["n"][f[3]], mesh["n"][f[4]], mesh["n"][f[5]]
                normal = norm(add(add(scale(N1, r1), scale(N2, r2)), scale(N3, r3)))          # triangle.go:29-31
            else:
                normal = norm(cross(sub(P2, P1), sub(P3, P1)))                                  # triangle.go:24-26
            dist = length(sub(o, ip))
            if best is None or dist < best[0]:
                best = (dist, ip, normal, mesh["m"][f[6]], fi)
        if best is None:
            return None
        return add(best[1], pos), best[2], best[3], best[4]

    # worker/shared/tracer/tracer.go:27-50
    def trace(self, o, d):
        nearest = None
        for k, (pos, mi) in enumerate(self.objs):
            r = self.object_intersection(pos, self.meshes[mi], o, d)
            if r is None:
                continue
            ip, n, mat, fi = r
            dist = length(sub(ip, self.cam_pos))             # distance to the CAMERA, also for shadow rays (:38)
            if nearest is None or dist < nearest[0]:
                nearest = (dist, ip, n, mat, fi, k)
        return nearest

    # tracer.go:53-77 + shared/colour/colour.go:38-50
    def phong(self, ip, n, mat):
        Ka, Kd, Ks, Ns = mat[0:3], mat[3:6], mat[6:9], mat[9]
        c = Ka
        lit = 0
        for li, (lpos, lcol) in enumerate(self.lights):
            ld = norm(sub(lpos, ip))
            sh = self.trace(add(ip, scale(ld, 0.0001)), ld)
            if sh is None or length(sub(lpos, ip)) < length(sub(sh[1], ip)):
                lit |= 1 << li
                refl = sub(scale(n, 2 * dot(ld, n)), ld)
                cd = norm(sub(self.cam_pos, ip))
                s = go_max(dot(ld, n), 0.0)
                diff = tuple(go_max(0.0, go_min(s * k, 1.0)) * l for k, l in zip(Kd, lcol))
                c = tuple(go_min(a + b, 1.0) for a, b in zip(c, diff))
                base = go_max(dot(refl, cd), 0.0)
                p = 1.0 if Ns == 0 else (base if Ns == 1 else go_pow_int(base, Ns))
                spec = tuple(go_max(0.0, go_min(p * k, 1.0)) * l for k, l in zip(Ks, lcol))
                c = tuple(go_min(a + b, 1.0) for a, b in zip(c, spec))
        return c, lit

    # tracer.go:15-22
    def pixel_to_point(self, i, j, W, H, tan_half):
        hw, hh = W // 2, H // 2
        phw = tan_half
        phh = phw * float(H) / float(W)
        io = scale(self.left, phw * (float(hw - i) - 0.5) / float(hw))
        jo = scale(self.up, phh * (float(hh - j) - 0.5) / float(hh))
        return add(add(add(self.cam_pos, self.fwd), io), jo)

    # tracer.go:81-91; tan_half = Go's math.Tan(Fov/2) (a frame constant, supplied by the caller)
    def trace_pixel(self, i, j, W, H, tan_half):
        pt = self.pixel_to_point(i, j, W, H, tan_half)
        r = self.trace(self.cam_pos, norm(sub(pt, self.cam_pos)))
        if r is None:
            return False, (0.0, 0.0, 0.0), -1, -1, 0
        _, ip, n, mat, fi, k = r
        c, lit = self.phong(ip, n, mat)
        return True, c, fi, k, lit


def quantise_rgb(c):                                 # colour.go:59-61  uint8(255 * c)
    return tuple(int(255 * v) & 0xff for v in c)


def quantise_rgba16_shr8(c):                         # colour.go:54-56 RGBA() then go-sdl2's >> 8
    return tuple((int(0xFFFF * v) & 0xffffffff) >> 8 for v in c)
