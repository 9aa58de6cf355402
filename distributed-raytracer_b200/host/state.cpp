// state.cpp -- see state.hpp.  Host translation units are compiled with -ffp-contract=off so that the
// float64 arithmetic below rounds exactly as the Go reference does (Go never fuses a*b+c on amd64).
#include "state.hpp"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../csrc/go_math.h"

namespace geom {

Vector Vector::Rotate(const Vector &b, double theta) const {   // vector.go:39-42 (Rodrigues)
    return Scale(std::cos(theta)).Add(b.Cross(*this).Scale(std::sin(theta))).Add(b.Scale(b.Dot(*this) * (1.0 - std::cos(theta))));
}
Vector Vector::Norm() const {                                  // vector.go:50-53
    double mag = std::sqrt(X * X + Y * Y + Z * Z);
    return { X / mag, Y / mag, Z / mag };
}
double Vector::Len() const { return std::sqrt(X * X + Y * Y + Z * Z); }   // vector.go:56-58
bool Vector::operator<(const Vector &o) const {
    // +0.0 is added so that -0.0 and +0.0 compare equal, like Go map keys (==)
    double a[3] = { X + 0.0, Y + 0.0, Z + 0.0 }, b[3] = { o.X + 0.0, o.Y + 0.0, o.Z + 0.0 };
    for (int k = 0; k < 3; k++) {
        if (a[k] < b[k]) return true;
        if (a[k] > b[k]) return false;
    }
    return false;
}

}  // namespace geom

namespace colour {
RGB NewRGB(uint8_t r, uint8_t g, uint8_t b) { return { double(r) / 255.0, double(g) / 255.0, double(b) / 255.0 }; }
static double clamp01(double v) { return std::fmax(0.0, std::fmin(v, 1.0)); }
RGB NewRGBFromFloats(float r, float g, float b) { return { clamp01(double(r)), clamp01(double(g)), clamp01(double(b)) }; }
}  // namespace colour

namespace state {

const geom::Vector GlobalUp = { 0.0, 1.0, 0.0 };

bool Material::operator<(const Material &o) const {
    const double a[10] = { Ka.r, Ka.g, Ka.b, Kd.r, Kd.g, Kd.b, Ks.r, Ks.g, Ks.b, Ns };
    const double b[10] = { o.Ka.r, o.Ka.g, o.Ka.b, o.Kd.r, o.Kd.g, o.Kd.b, o.Ks.r, o.Ks.g, o.Ks.b, o.Ns };
    for (int k = 0; k < 10; k++) {
        if (a[k] < b[k]) return true;
        if (a[k] > b[k]) return false;
    }
    return false;
}

// util.go:11 relativePath
static std::string relativePath(const std::string &original, const std::string &other) {
    size_t cut = original.find_last_of("/\\");
    std::string dir = cut == std::string::npos ? "" : original.substr(0, cut + 1);
    size_t s = other.find_first_not_of("/\\");
    return dir + (s == std::string::npos ? "" : other.substr(s));
}

static bool file_exists(const std::string &p) { std::ifstream f(p); return f.good(); }

// ---- gwob-equivalent OBJ / MTL reading (float32 precision; tri + quad faces; quad -> 0 1 2, 2 3 0) ----
struct MtlEntry { float Ka[3] = { 0, 0, 0 }, Kd[3] = { 0, 0, 0 }, Ks[3] = { 0, 0, 0 }; float Ns = 0; };

static std::vector<std::string> fields(const std::string &line) {
    std::vector<std::string> out;
    std::istringstream is(line);
    std::string t;
    while (is >> t) out.push_back(t);
    return out;
}
static float parse_f32(const std::string &s) { return std::strtof(s.c_str(), nullptr); }   // strconv.ParseFloat(s, 32)

static std::map<std::string, MtlEntry> ReadMaterialLibFromFile(const std::string &path) {
    std::ifstream f(path);
    if (!f.good()) throw std::runtime_error("cannot open material library " + path);
    std::map<std::string, MtlEntry> lib;
    MtlEntry *cur = nullptr;
    std::string line;
    while (std::getline(f, line)) {
        auto t = fields(line);
        if (t.empty() || t[0][0] == '#') continue;
        if (t[0] == "newmtl") cur = &lib[t.size() > 1 ? t[1] : ""];
        else if (cur && t.size() >= 4 && (t[0] == "Ka" || t[0] == "Kd" || t[0] == "Ks")) {
            float *dst = t[0] == "Ka" ? cur->Ka : (t[0] == "Kd" ? cur->Kd : cur->Ks);
            for (int k = 0; k < 3; k++) dst[k] = parse_f32(t[1 + k]);
        } else if (cur && t.size() >= 2 && t[0] == "Ns") cur->Ns = parse_f32(t[1]);
    }
    return lib;
}

struct Corner { int v = -1, n = -1; };
struct ObjGroup { std::string usemtl; std::vector<Corner> corners; };   // 3 corners per triangle

std::shared_ptr<Mesh> MeshFromFile(const std::string &path) {
    std::ifstream f(path);
    if (!f.good()) throw std::runtime_error("cannot open model " + path);
    std::vector<float> vs, vns;
    std::vector<ObjGroup> groups(1);
    std::string mtllib, line;
    bool normFound = false;
    while (std::getline(f, line)) {
        auto t = fields(line);
        if (t.empty() || t[0][0] == '#') continue;
        const std::string &k = t[0];
        if (k == "v" && t.size() >= 4) { for (int c = 0; c < 3; c++) vs.push_back(parse_f32(t[1 + c])); }
        else if (k == "vn" && t.size() >= 4) { for (int c = 0; c < 3; c++) vns.push_back(parse_f32(t[1 + c])); }
        else if (k == "mtllib" && t.size() >= 2) mtllib = t[1];
        else if (k == "o" || k == "g") {
            if (!groups.back().corners.empty()) { ObjGroup g; g.usemtl = groups.back().usemtl; groups.push_back(g); }
        } else if (k == "usemtl") {
            std::string name = t.size() > 1 ? t[1] : "";
            if (!groups.back().corners.empty()) { ObjGroup g; g.usemtl = name; groups.push_back(g); }
            else groups.back().usemtl = name;
        } else if (k == "f") {
            std::vector<Corner> cs;
            for (size_t q = 1; q < t.size(); q++) {
                Corner c;
                const std::string &tok = t[q];
                size_t s1 = tok.find('/');
                int vi = std::atoi(tok.substr(0, s1).c_str());
                c.v = vi > 0 ? vi - 1 : int(vs.size() / 3) + vi;
                if (s1 != std::string::npos) {
                    size_t s2 = tok.find('/', s1 + 1);
                    if (s2 != std::string::npos && s2 + 1 < tok.size()) {
                        int ni = std::atoi(tok.substr(s2 + 1).c_str());
                        c.n = ni > 0 ? ni - 1 : int(vns.size() / 3) + ni;
                        normFound = true;
                    }
                }
                if (c.v < 0 || size_t(c.v) >= vs.size() / 3) throw std::runtime_error("OBJ vertex index out of range in " + path);
                if (c.n >= 0 && size_t(c.n) >= vns.size() / 3) throw std::runtime_error("OBJ normal index out of range in " + path);
                cs.push_back(c);
            }
            auto &dst = groups.back().corners;
            if (cs.size() == 3) { dst.insert(dst.end(), { cs[0], cs[1], cs[2] }); }
            else if (cs.size() == 4) { dst.insert(dst.end(), { cs[0], cs[1], cs[2], cs[2], cs[3], cs[0] }); }
            else throw std::runtime_error("OBJ face with unsupported vertex count in " + path);
        }
    }
    std::map<std::string, MtlEntry> lib;
    if (!mtllib.empty()) {                                                           // mesh.go:120-129
        std::string rel = relativePath(path, mtllib);
        lib = ReadMaterialLibFromFile(file_exists(rel) ? rel : mtllib);
    }
    auto mesh = std::make_shared<Mesh>();
    std::map<geom::Vector, unsigned> vertexMap, vertexNormalMap;                     // mesh.go:146-148
    std::map<Material, unsigned> materialMap;
    for (const ObjGroup &g : groups) {
        if (g.corners.empty()) continue;
        Material mat;                                                                // mesh.go:151 default material
        mat.Ka = colour::NewRGB(0x10, 0x10, 0x10); mat.Kd = colour::NewRGB(0xFF, 0xFF, 0xFF); mat.Ks = colour::NewRGB(0, 0, 0); mat.Ns = 0.0;
        auto it = lib.find(g.usemtl);
        if (it != lib.end()) {                                                       // mesh.go:152-154
            const MtlEntry &e = it->second;
            mat.Ka = colour::NewRGBFromFloats(e.Ka[0], e.Ka[1], e.Ka[2]);
            mat.Kd = colour::NewRGBFromFloats(e.Kd[0], e.Kd[1], e.Kd[2]);
            mat.Ks = colour::NewRGBFromFloats(e.Ks[0], e.Ks[1], e.Ks[2]);
            mat.Ns = double(e.Ns);
        }
        unsigned matIndex;
        auto mi = materialMap.find(mat);                                             // mesh.go:157-163
        if (mi == materialMap.end()) { matIndex = unsigned(mesh->materials.size()); mesh->materials.push_back(mat); materialMap[mat] = matIndex; }
        else matIndex = mi->second;
        for (size_t fi = 0; fi + 2 < g.corners.size(); fi += 3) {                    // mesh.go:166
            face fc;
            fc.mat = matIndex;
            for (int v = 0; v < 3; v++) {
                const Corner &c = g.corners[fi + v];
                geom::Vector vv = { double(vs[3 * c.v]), double(vs[3 * c.v + 1]), double(vs[3 * c.v + 2]) };   // Coord64, mesh.go:176
                auto vi = vertexMap.find(vv);                                        // mesh.go:182-188
                if (vi != vertexMap.end()) fc.verts[v] = vi->second;
                else { fc.verts[v] = unsigned(mesh->vertices.size()); vertexMap[vv] = fc.verts[v]; mesh->vertices.push_back(vv); }
                if (normFound) {                                                     // mesh.go:191-203
                    geom::Vector nn = { 0, 0, 0 };
                    if (c.n >= 0) nn = { double(vns[3 * c.n]), double(vns[3 * c.n + 1]), double(vns[3 * c.n + 2]) };
                    auto ni = vertexNormalMap.find(nn);
                    if (ni != vertexNormalMap.end()) fc.vertNorms[v] = ni->second;
                    else {
                        fc.vertNorms[v] = unsigned(mesh->vertexNormals.size());
                        vertexNormalMap[nn] = fc.vertNorms[v];
                        mesh->vertexNormals.push_back(nn.Norm());
                    }
                }
            }
            mesh->faces.push_back(fc);                                               // mesh.go:208 (Insert)
        }
    }
    return mesh;
}

Camera NewCamera(const geom::Vector &pos, const geom::Vector &dir, double fov) {   // camera.go:35-44
    if (dir.Cross(GlobalUp).Zero()) throw std::runtime_error("Camera dir is parallel to global up.");
    Camera c;
    c.Pos = pos;
    c.forward = dir.Norm();
    c.left = dir.Cross(GlobalUp).Norm();
    c.up = c.left.Cross(c.forward);
    c.Fov = fov;
    return c;
}

void Camera::Move(double distance, bool fwd, bool backward, bool leftward, bool rightward, bool upward, bool downward) {  // camera.go:62-92
    geom::Vector moveDir = { 0, 0, 0 };
    if (fwd != backward) moveDir = fwd ? moveDir.Add(forward) : moveDir.Sub(forward);
    if (leftward != rightward) moveDir = leftward ? moveDir.Add(left) : moveDir.Sub(left);
    if (upward != downward) moveDir = upward ? moveDir.Add(up) : moveDir.Sub(up);
    if (!moveDir.Zero()) Pos = Pos.Add(moveDir.Norm().Scale(distance));
}

void Camera::Yaw(double theta) {                                                    // camera.go:130-146
    if (std::fmod(theta, 2.0 * M_PI) != 0.0) {
        forward = forward.Rotate(up, theta).Norm();
        // camera.go:135-137 nudges forward randomly if it became parallel to GlobalUp; a deterministic
        // nudge along +x is used here (the reference seeds math/rand from the wall clock)
        if (forward.Cross(GlobalUp).Zero()) forward = forward.Add({ 0.0001, 0.0, 0.0 });
        left = forward.Cross(GlobalUp).Norm();
        up = left.Cross(forward).Norm();
    }
}

void Camera::Pitch(double theta) {                                                  // camera.go:149-154
    if (std::fmod(theta, 2.0 * M_PI) != 0.0) {
        forward = forward.Rotate(left, theta).Norm();
        up = left.Cross(forward).Norm();
    }
}

// ---- minimal JSON reader (objects, arrays, numbers, strings, literals) for scene.json ----
struct JVal {
    enum Kind { Null, Num, Str, Arr, Obj, Bool } kind = Null;
    double num = 0;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal *get(const std::string &k) const {
        for (auto &kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};
struct JParser {
    const std::string &s;
    size_t p = 0;
    explicit JParser(const std::string &src) : s(src) {}
    void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) p++; }
    [[noreturn]] void bad(const char *m) { throw std::runtime_error(std::string("scene JSON: ") + m); }
    JVal parse() {
        ws();
        if (p >= s.size()) bad("unexpected end");
        char c = s[p];
        JVal v;
        if (c == '{') {
            v.kind = JVal::Obj; p++; ws();
            if (s[p] == '}') { p++; return v; }
            for (;;) {
                ws();
                JVal k = parse();
                if (k.kind != JVal::Str) bad("object key must be a string");
                ws();
                if (s[p++] != ':') bad("expected ':'");
                v.obj.emplace_back(k.str, parse());
                ws();
                if (s[p] == ',') { p++; continue; }
                if (s[p] == '}') { p++; break; }
                bad("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = JVal::Arr; p++; ws();
            if (s[p] == ']') { p++; return v; }
            for (;;) {
                v.arr.push_back(parse());
                ws();
                if (s[p] == ',') { p++; continue; }
                if (s[p] == ']') { p++; break; }
                bad("expected ',' or ']'");
            }
        } else if (c == '"') {
            v.kind = JVal::Str; p++;
            while (p < s.size() && s[p] != '"') { if (s[p] == '\\' && p + 1 < s.size()) p++; v.str.push_back(s[p++]); }
            if (p >= s.size()) bad("unterminated string");
            p++;
        } else if (c == 't' || c == 'f' || c == 'n') {
            v.kind = c == 'n' ? JVal::Null : JVal::Bool;
            while (p < s.size() && std::isalpha((unsigned char)s[p])) p++;
        } else {
            char *end = nullptr;
            v.kind = JVal::Num;
            v.num = std::strtod(s.c_str() + p, &end);
            if (end == s.c_str() + p) bad("bad number");
            p = size_t(end - s.c_str());
        }
        return v;
    }
};
static double jnum(const JVal *o, const char *k) {
    const JVal *v = o ? o->get(k) : nullptr;
    return v && v->kind == JVal::Num ? v->num : 0.0;
}
static geom::Vector jvec(const JVal *o) { return { jnum(o, "x"), jnum(o, "y"), jnum(o, "z") }; }

rt_obj_mesh_desc ObjTextMesh::desc() const {
    rt_obj_mesh_desc d;
    d.obj_text = text.data(); d.obj_bytes = text.size();
    d.materials = materials.data(); d.n_materials = uint32_t(materials.size() / 10);
    d.usemtl_offset = usemtl_off.empty() ? nullptr : usemtl_off.data();
    d.usemtl_material = usemtl_mat.empty() ? nullptr : usemtl_mat.data();
    d.n_usemtl = uint32_t(usemtl_off.size());
    d.default_material = default_material;
    return d;
}

static void material_doubles(const Material &m, double out[10]) {
    const double a[10] = { m.Ka.r, m.Ka.g, m.Ka.b, m.Kd.r, m.Kd.g, m.Kd.b, m.Ks.r, m.Ks.g, m.Ks.b, m.Ns };
    for (int k = 0; k < 10; k++) out[k] = a[k];
}

ObjTextMesh ObjTextFromFile(const std::string &path) {
    ObjTextMesh o;
    {
        std::ifstream f(path, std::ios::binary);
        if (!f.good()) throw std::runtime_error("cannot open model " + path);
        std::stringstream ss;
        ss << f.rdbuf();
        o.text = ss.str();
    }
    // one pass over the line starts: only `mtllib` and `usemtl` lines are looked at
    std::string mtllib;
    std::vector<std::pair<uint64_t, std::string>> uses;
    const char *base = o.text.data(), *end = base + o.text.size();
    for (const char *p = base; p < end;) {
        const char *nl = (const char *)memchr(p, '\n', size_t(end - p));
        const char *le = nl ? nl : end;
        const char *q = p;
        while (q < le && (*q == ' ' || *q == '\t')) q++;
        if (le - q > 6 && (!memcmp(q, "usemtl", 6) || !memcmp(q, "mtllib", 6)) && (q[6] == ' ' || q[6] == '\t')) {
            auto t = fields(std::string(q, le));
            std::string name = t.size() > 1 ? t[1] : "";
            if (q[0] == 'u') uses.emplace_back(uint64_t(p - base), name); else mtllib = name;
        }
        p = nl ? nl + 1 : end;
    }
    std::map<std::string, MtlEntry> lib;
    if (!mtllib.empty()) {                                                           // mesh.go:120-129
        std::string rel = relativePath(path, mtllib);
        lib = ReadMaterialLibFromFile(file_exists(rel) ? rel : mtllib);
    }
    std::map<Material, unsigned> materialMap;
    auto index_of = [&](const std::string &name) {
        Material mat;                                                                // mesh.go:151 default material
        mat.Ka = colour::NewRGB(0x10, 0x10, 0x10); mat.Kd = colour::NewRGB(0xFF, 0xFF, 0xFF); mat.Ks = colour::NewRGB(0, 0, 0); mat.Ns = 0.0;
        auto it = lib.find(name);
        if (it != lib.end()) {                                                       // mesh.go:152-154
            const MtlEntry &e = it->second;
            mat.Ka = colour::NewRGBFromFloats(e.Ka[0], e.Ka[1], e.Ka[2]);
            mat.Kd = colour::NewRGBFromFloats(e.Kd[0], e.Kd[1], e.Kd[2]);
            mat.Ks = colour::NewRGBFromFloats(e.Ks[0], e.Ks[1], e.Ks[2]);
            mat.Ns = double(e.Ns);
        }
        auto mi = materialMap.find(mat);                                             // mesh.go:157-163
        if (mi != materialMap.end()) return mi->second;
        unsigned k = unsigned(o.materials.size() / 10);
        double d[10];
        material_doubles(mat, d);
        o.materials.insert(o.materials.end(), d, d + 10);
        materialMap[mat] = k;
        return k;
    };
    o.default_material = index_of("");                                               // faces above the first usemtl line
    for (auto &u : uses) { o.usemtl_off.push_back(u.first); o.usemtl_mat.push_back(index_of(u.second)); }
    return o;
}

int GpuSceneFromFile(const std::string &path, int device, uint32_t precision, rt_scene **out, EnvMutables &em, std::string &err) {
    *out = nullptr;
    try {
        std::ifstream f(path);
        if (!f.good()) throw std::runtime_error("cannot open scene " + path);
        std::stringstream ss;
        ss << f.rdbuf();
        std::string src = ss.str();
        JParser jp(src);
        JVal root = jp.parse();
        std::vector<ObjTextMesh> texts;
        std::map<std::string, uint32_t> index;
        std::vector<double> obj_pos;
        std::vector<uint32_t> obj_mesh;
        em = EnvMutables();
        const JVal *objs = root.get("objs");
        if (objs)
            for (size_t i = 0; i < objs->arr.size(); i++) {                         // environment.go:190-216
                const JVal &o = objs->arr[i];
                const JVal *model = o.get("model");
                std::string m = model ? model->str : "";
                auto it = index.find(m);
                if (it == index.end()) {
                    std::string rel = relativePath(path, m);
                    texts.push_back(ObjTextFromFile(file_exists(rel) ? rel : m));
                    it = index.emplace(m, uint32_t(texts.size() - 1)).first;
                }
                Object ob;
                ob.Pos = jvec(o.get("pos")); ob.id = unsigned(i + 1);
                em.Objs.push_back(ob);
                obj_pos.insert(obj_pos.end(), { ob.Pos.X, ob.Pos.Y, ob.Pos.Z });
                obj_mesh.push_back(it->second);
            }
        const JVal *lights = root.get("lights");
        if (lights)
            for (const JVal &l : lights->arr) {                                      // environment.go:220-225
                Light li;
                li.Pos = jvec(l.get("pos"));
                const JVal *c = l.get("col");
                li.Col = colour::NewRGB(uint8_t(jnum(c, "r")), uint8_t(jnum(c, "g")), uint8_t(jnum(c, "b")));
                em.Lights.push_back(li);
            }
        const JVal *cam = root.get("cam");
        em.Cam = NewCamera(jvec(cam ? cam->get("pos") : nullptr), jvec(cam ? cam->get("dir") : nullptr), jnum(cam, "fov"));
        std::vector<rt_obj_mesh_desc> descs;
        for (const ObjTextMesh &t : texts) descs.push_back(t.desc());
        int rc = rt_scene_create_obj(descs.data(), uint32_t(descs.size()), obj_pos.data(), obj_mesh.data(), uint32_t(obj_mesh.size()), device, precision, out);
        if (rc != RT_OK) err = rt_last_error(nullptr);
        return rc;
    } catch (const std::exception &e) {
        err = e.what();
        return RT_ERR_INVALID;
    }
}

Environment EnvironmentFromFile(const std::string &path) {                          // environment.go:162-234
    std::ifstream f(path);
    if (!f.good()) throw std::runtime_error("cannot open scene " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string src = ss.str();
    JParser jp(src);
    JVal root = jp.parse();
    Environment env;
    const JVal *objs = root.get("objs");
    if (objs)
        for (size_t i = 0; i < objs->arr.size(); i++) {
            const JVal &o = objs->arr[i];
            const JVal *model = o.get("model");
            std::string m = model ? model->str : "";
            auto it = env.meshes.find(m);
            std::shared_ptr<Mesh> mesh;
            if (it == env.meshes.end()) {                                            // environment.go:194-205
                std::string rel = relativePath(path, m);
                mesh = MeshFromFile(file_exists(rel) ? rel : m);
                env.meshes[m] = mesh;
            } else mesh = it->second;
            env.paths[unsigned(i + 1)] = m;                                          // environment.go:209
            Object ob;
            ob.Pos = jvec(o.get("pos")); ob.id = unsigned(i + 1); ob.mesh = mesh;
            env.mutableState.Objs.push_back(ob);                                     // environment.go:212
        }
    const JVal *lights = root.get("lights");
    if (lights)
        for (const JVal &l : lights->arr) {                                          // environment.go:220-225
            Light li;
            li.Pos = jvec(l.get("pos"));
            const JVal *c = l.get("col");
            li.Col = colour::NewRGB(uint8_t(jnum(c, "r")), uint8_t(jnum(c, "g")), uint8_t(jnum(c, "b")));
            env.mutableState.Lights.push_back(li);
        }
    const JVal *cam = root.get("cam");
    env.mutableState.Cam = NewCamera(jvec(cam ? cam->get("pos") : nullptr), jvec(cam ? cam->get("dir") : nullptr), jnum(cam, "fov"));  // :228
    return env;
}

void Environment::Flatten(FlatScene &out, int device, uint32_t precision) const {
    out = FlatScene();
    std::map<const Mesh *, uint32_t> index;
    for (const Object &o : mutableState.Objs) {
        const Mesh *m = o.mesh.get();
        if (!m) throw std::runtime_error("object without a mesh cannot be flattened");
        auto it = index.find(m);
        uint32_t mi;
        if (it == index.end()) {
            mi = uint32_t(out.meshes.size());
            index[m] = mi;
            out.meshes.emplace_back();
            FlatScene::FlatMesh &fm = out.meshes.back();
            for (const geom::Vector &v : m->vertices) { fm.verts.push_back(v.X); fm.verts.push_back(v.Y); fm.verts.push_back(v.Z); }
            for (const geom::Vector &v : m->vertexNormals) { fm.normals.push_back(v.X); fm.normals.push_back(v.Y); fm.normals.push_back(v.Z); }
            for (const face &fc : m->faces) {
                for (int k = 0; k < 3; k++) fm.faces.push_back(fc.verts[k]);
                for (int k = 0; k < 3; k++) fm.faces.push_back(fc.vertNorms[k]);
                fm.faces.push_back(fc.mat);
            }
            for (const Material &mt : m->materials) {
                const double a[10] = { mt.Ka.r, mt.Ka.g, mt.Ka.b, mt.Kd.r, mt.Kd.g, mt.Kd.b, mt.Ks.r, mt.Ks.g, mt.Ks.b, mt.Ns };
                fm.materials.insert(fm.materials.end(), a, a + 10);
            }
        } else mi = it->second;
        out.obj_pos.push_back(o.Pos.X); out.obj_pos.push_back(o.Pos.Y); out.obj_pos.push_back(o.Pos.Z);
        out.obj_mesh.push_back(mi);
    }
    for (FlatScene::FlatMesh &fm : out.meshes) {
        rt_mesh_desc d;
        d.verts = fm.verts.data(); d.n_verts = uint32_t(fm.verts.size() / 3);
        d.normals = fm.normals.empty() ? nullptr : fm.normals.data(); d.n_normals = uint32_t(fm.normals.size() / 3);
        d.faces = fm.faces.data(); d.n_faces = uint32_t(fm.faces.size() / 7);
        d.materials = fm.materials.data(); d.n_materials = uint32_t(fm.materials.size() / 10);
        out.mesh_descs.push_back(d);
    }
    out.desc.meshes = out.mesh_descs.data(); out.desc.n_meshes = uint32_t(out.mesh_descs.size());
    out.desc.obj_pos = out.obj_pos.data(); out.desc.obj_mesh = out.obj_mesh.data(); out.desc.n_objects = uint32_t(out.obj_mesh.size());
    out.desc.device = device; out.desc.precision = precision; out.desc.flags = RT_SCENE_DEFAULT;
}

void FrameState(const EnvMutables &em, rt_camera &cam, std::vector<rt_light> &lights, std::vector<double> &obj_pos) {
    const Camera &c = em.Cam;
    const geom::Vector v[4] = { c.Pos, c.Forward(), c.Left(), c.Up() };
    double *dst[4] = { cam.pos, cam.forward, cam.left, cam.up };
    for (int k = 0; k < 4; k++) { dst[k][0] = v[k].X; dst[k][1] = v[k].Y; dst[k][2] = v[k].Z; }
    cam.fov = c.Fov;
    cam.tan_half_fov = rt_go_tan(c.Fov / 2.0);               // tracer.go:17, Go's own math.Tan algorithm
    lights.clear();
    for (const Light &l : em.Lights) {
        rt_light r;
        r.pos[0] = l.Pos.X; r.pos[1] = l.Pos.Y; r.pos[2] = l.Pos.Z;
        r.col[0] = l.Col.r; r.col[1] = l.Col.g; r.col[2] = l.Col.b;
        lights.push_back(r);
    }
    obj_pos.clear();
    for (const Object &o : em.Objs) { obj_pos.push_back(o.Pos.X); obj_pos.push_back(o.Pos.Y); obj_pos.push_back(o.Pos.Z); }
}

}  // namespace state
