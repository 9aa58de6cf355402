// state.hpp -- C++ mirror of the reference's Go host-side scene model for the render path
// (the Go toolchain is absent from this image, so the host side above the C ABI is C++; the cgo
// binding a maintainer would use instead is in INTEGRATION.md).
//
// Mirrors, with the same names and argument meaning (paths under /root/reference/):
//   shared/geom/vector.go          geom::Vector
//   shared/colour/colour.go:16-35  colour::RGB, NewRGB, NewRGBFromFloats
//   shared/state/mesh.go:21-27,94-106,109-213   state::face, Material, Mesh, MeshFromFile
//   shared/state/object.go:17-22   state::Object
//   shared/state/light.go:10-13    state::Light
//   shared/state/camera.go:20-154  state::Camera (NewCamera, Forward/Left/Up, Move, Yaw, Pitch)
//   shared/state/environment.go:65-69,162-234,271   state::EnvMutables, Environment, EnvironmentFromFile, Mutable()
// plus the one addition the GPU backend needs: Environment::Flatten(), the SoA
// flattening that fills an rt_scene_desc.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rt_b200.h"

namespace geom {
struct Vector {
    double X = 0, Y = 0, Z = 0;
    Vector Add(const Vector &b) const { return { X + b.X, Y + b.Y, Z + b.Z }; }
    Vector Sub(const Vector &b) const { return { X - b.X, Y - b.Y, Z - b.Z }; }
    Vector Scale(double s) const { return { s * X, s * Y, s * Z }; }
    double Dot(const Vector &b) const { return X * b.X + Y * b.Y + Z * b.Z; }
    Vector Cross(const Vector &b) const { return { Y * b.Z - Z * b.Y, Z * b.X - X * b.Z, X * b.Y - Y * b.X }; }
    Vector Rotate(const Vector &b, double theta) const;
    bool Zero() const { return X == 0.0 && Y == 0.0 && Z == 0.0; }
    Vector Norm() const;
    double Len() const;
    bool operator<(const Vector &o) const;   // map key (value identity; -0 == +0 as in Go map keys)
};
}  // namespace geom

namespace colour {
struct RGB {
    double r = 0, g = 0, b = 0;
};
RGB NewRGB(uint8_t r, uint8_t g, uint8_t b);                 // colour.go:28-30
RGB NewRGBFromFloats(float r, float g, float b);             // colour.go:33-35
}  // namespace colour

namespace state {

extern const geom::Vector GlobalUp;                          // environment.go:22

struct Material {                                            // mesh.go:94-97
    colour::RGB Ka, Kd, Ks;
    double Ns = 0;
    bool operator<(const Material &o) const;
};

struct face {                                                // mesh.go:21-27
    unsigned verts[3] = { 0, 0, 0 };
    unsigned vertNorms[3] = { 0, 0, 0 };
    unsigned mat = 0;
};

class Mesh {                                                 // mesh.go:100-106
  public:
    std::vector<geom::Vector> vertices;
    std::vector<geom::Vector> vertexNormals;
    std::vector<face> faces;          // enumeration order (the reference keeps them only in an R-tree)
    std::vector<Material> materials;
};

std::shared_ptr<Mesh> MeshFromFile(const std::string &path);   // mesh.go:109-213 (throws std::runtime_error)

struct Object {                                              // object.go:17-22
    geom::Vector Pos;
    unsigned id = 0;
    std::shared_ptr<Mesh> mesh;
};

struct Light {                                               // light.go:10-13
    geom::Vector Pos;
    colour::RGB Col;
};

class Camera {                                               // camera.go:20-24
  public:
    geom::Vector Pos;
    double Fov = 0;
    const geom::Vector &Forward() const { return forward; }
    const geom::Vector &Left() const { return left; }
    const geom::Vector &Up() const { return up; }
    void Move(double distance, bool forward, bool backward, bool leftward, bool rightward, bool upward, bool downward);  // :62
    void Yaw(double theta);                                  // :130
    void Pitch(double theta);                                // :149
    friend Camera NewCamera(const geom::Vector &pos, const geom::Vector &dir, double fov);
    void SetBasis(const geom::Vector &f, const geom::Vector &l, const geom::Vector &u) { forward = f; left = l; up = u; }

  private:
    geom::Vector forward, left, up;
};
Camera NewCamera(const geom::Vector &pos, const geom::Vector &dir, double fov);   // camera.go:35-44 (throws)

struct EnvMutables {                                         // environment.go:65-69
    std::vector<Object> Objs;
    std::vector<Light> Lights;
    Camera Cam;
};

// Flat SoA image of an Environment, laid out exactly as include/rt_b200.h expects.
struct FlatScene {
    struct FlatMesh {
        std::vector<double> verts, normals, materials;
        std::vector<uint32_t> faces;
    };
    std::vector<FlatMesh> meshes;
    std::vector<rt_mesh_desc> mesh_descs;
    std::vector<double> obj_pos;
    std::vector<uint32_t> obj_mesh;
    rt_scene_desc desc;
};

class Environment {                                          // environment.go:149-153
  public:
    std::map<std::string, std::shared_ptr<Mesh>> meshes;     // immutable: model path -> mesh
    std::map<unsigned, std::string> paths;                   // immutable: object id -> model path
    EnvMutables mutableState;
    EnvMutables &Mutable() { return mutableState; }          // environment.go:271
    // SoA flattening for rt_scene_create (meshes in first-use order)
    void Flatten(FlatScene &out, int device, uint32_t precision) const;
};

Environment EnvironmentFromFile(const std::string &path);    // environment.go:162-234 (throws)

// a model as the TEXT of its OBJ file + the small tables the GPU ingest needs (rt_obj_mesh_desc): the host only reads
// the bytes, finds the mtllib / usemtl lines and parses the MTL file -- no tokenizer pass over the geometry, no value-keyed maps.
struct ObjTextMesh {
    std::string text;
    std::vector<double> materials;            // 10 per material, first-use order, de-duplicated by value (mesh.go:151-163)
    std::vector<uint64_t> usemtl_off;
    std::vector<uint32_t> usemtl_mat;
    uint32_t default_material = 0;
    rt_obj_mesh_desc desc() const;
};
ObjTextMesh ObjTextFromFile(const std::string &path);        // throws
// register() of a GPU worker straight from the scene file: every distinct model goes to the device as OBJ text
// (rt_scene_create_obj); `em` receives objects, lights and camera exactly as EnvironmentFromFile fills them.
int GpuSceneFromFile(const std::string &scene_path, int device, uint32_t precision, rt_scene **out, EnvMutables &em, std::string &err);

// frame state for rt_frame_set / rt_bulk_trace from the mutable part
void FrameState(const EnvMutables &em, rt_camera &cam, std::vector<rt_light> &lights, std::vector<double> &obj_pos);

}  // namespace state
