// Host side of the drop-in Raytracer class (see raytracer.h).  Compiled with
// -ffp-contract=off: the few float computations that stay on the host (model matrices, vertex
// transforms, camera basis, view-matrix inverse) must round exactly like the reference's.
#include "raytracer.h"
#include "json_min.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

namespace {

using V3 = Raytracer::Vector3;
using M4 = Raytracer::Matrix;

constexpr double kPI = 3.14159265;   // h:11

inline V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
inline V3 sub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross3(V3 a, V3 b) { return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline V3 normalized(V3 a) {                     // h:109-116: sqrt, then three divisions, skipped at length 0
    float len = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    if (len > 0) { a.x /= len; a.y /= len; a.z /= len; }
    return a;
}
inline float toRadian(float degrees) { return degrees * (kPI / 180); }   // h:581-583 (double product, float result)

void identity(M4& a) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) a.m[i][j] = (i == j) ? 1.0f : 0.0f; }
M4 mul(const M4& a, const M4& b) {               // h:179-190: accumulate from 0 in k order
    M4 r;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            r.m[i][j] = 0;
            for (int k = 0; k < 4; ++k) r.m[i][j] += a.m[i][k] * b.m[k][j];
        }
    return r;
}
V3 transformPoint(const M4& M, V3 p) {           // h:234-248
    float x = M.m[0][0] * p.x + M.m[0][1] * p.y + M.m[0][2] * p.z + M.m[0][3];
    float y = M.m[1][0] * p.x + M.m[1][1] * p.y + M.m[1][2] * p.z + M.m[1][3];
    float z = M.m[2][0] * p.x + M.m[2][1] * p.y + M.m[2][2] * p.z + M.m[2][3];
    float w = M.m[3][0] * p.x + M.m[3][1] * p.y + M.m[3][2] * p.z + M.m[3][3];
    if (w != 1.0f) { x /= w; y /= w; z /= w; }
    return mk(x, y, z);
}
// 3x3 minor expansion on the upper-left block of a scratch matrix (h:251-255)
float minor3(const float s[3][3]) {
    return s[0][0] * (s[1][1] * s[2][2] - s[1][2] * s[2][1]) -
           s[0][1] * (s[1][0] * s[2][2] - s[1][2] * s[2][0]) +
           s[0][2] * (s[1][0] * s[2][1] - s[1][1] * s[2][0]);
}
// Matrix::Inverse (h:354-370) = Adjoint (h:276-296) / Determinant (h:257-274), same op order
int inverse4(const M4& a, M4& out) {
    float det = 0;
    for (int i = 0; i < 4; i++) {
        float s[3][3];
        for (int j = 1; j < 4; j++) {
            int c = 0;
            for (int k = 0; k < 4; k++) { if (k == i) continue; s[j - 1][c++] = a.m[j][k]; }
        }
        det += (i % 2 == 0 ? 1 : -1) * a.m[0][i] * minor3(s);
    }
    if (std::fabs(det) < 1e-10) return RT_FAILURE;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s[3][3];
            int si = 0;
            for (int k = 0; k < 4; k++) {
                if (k == i) continue;
                int sj = 0;
                for (int l = 0; l < 4; l++) { if (l == j) continue; s[si][sj++] = a.m[k][l]; }
                si++;
            }
            float cof = minor3(s);
            if ((i + j) % 2 != 0) cof = -cof;
            out.m[j][i] = cof / det;      // adjoint is the transposed cofactor matrix
        }
    return RT_SUCCESS;
}

bool readFile(const std::string& path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return false;
    std::stringstream ss; ss << f.rdbuf();
    out = ss.str();
    return true;
}

V3 vec3Of(const jsonmin::Value& a) { return mk(a.at(0).as_float(), a.at(1).as_float(), a.at(2).as_float()); }

}  // namespace

Raytracer::Raytracer(int width, int height) : mWidth(width), mHeight(height) {
    // cpp:781-788: Display{xRes,yRes}, frameBuffer = new Pixel[w*h], fov = 60.  The reference's
    // mGenerator stays default-seeded (cpp:787 shadows the member, Q1); the device replays
    // that stream (RT580_RNG_REFERENCE_LCG).
    if (width > 0 && height > 0) mFrameBuffer.resize((size_t)width * height);
}

Raytracer::~Raytracer() {
    if (mCtx) rt580_destroy(mCtx);
    for (rt580_context* c : mPeers) if (c) rt580_destroy(c);
    delete mScene;
}

// ---- binary mesh cache -------------------------------------------------------------------------------
// file = header { magic "RT580MSH", version, json size, FNV-1a 64 of the json bytes, type, radius, triangle count } + triangles
namespace {
struct MeshCacheHeader { char magic[8]; uint32_t version; uint32_t type; uint64_t json_size; uint64_t json_hash; float radius; uint32_t pad; uint64_t n_tris; };
uint64_t fnv1a64(const std::string& s) {
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
    return h;
}
}  // namespace

bool Raytracer::LoadMeshFromCache(const std::string& path, const std::string& jsonText, Mesh& mesh) {
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) return false;
    MeshCacheHeader h;
    if (!f.read(reinterpret_cast<char*>(&h), sizeof h)) return false;
    if (memcmp(h.magic, "RT580MSH", 8) != 0 || h.version != 1u || h.json_size != (uint64_t)jsonText.size() || h.json_hash != fnv1a64(jsonText)) return false;
    if (h.type > 1u || h.n_tris > (1ull << 32)) return false;
    static_assert(sizeof(Triangle) == 18 * sizeof(float), "Triangle is 18 packed floats");
    mesh.type = h.type == 0u ? Mesh::RT_POLYGON : Mesh::RT_SPHERE;
    mesh.radius = h.radius;
    mesh.triangles.resize((size_t)h.n_tris);
    if (h.n_tris && !f.read(reinterpret_cast<char*>(mesh.triangles.data()), (std::streamsize)(h.n_tris * sizeof(Triangle)))) return false;
    char extra;
    return !f.read(&extra, 1);               // nothing may follow
}

void Raytracer::StoreMeshInCache(const std::string& path, const std::string& jsonText, const Mesh& mesh) {
    const std::string tmp = path + ".tmp";
    {
        std::ofstream f(tmp, std::ios::binary);
        if (!f.is_open()) return;            // a read-only cache directory is not an error
        MeshCacheHeader h;
        memset(&h, 0, sizeof h);
        memcpy(h.magic, "RT580MSH", 8);
        h.version = 1u; h.type = mesh.type == Mesh::RT_POLYGON ? 0u : 1u;
        h.json_size = (uint64_t)jsonText.size(); h.json_hash = fnv1a64(jsonText);
        h.radius = mesh.radius; h.n_tris = (uint64_t)mesh.triangles.size();
        f.write(reinterpret_cast<const char*>(&h), sizeof h);
        if (!mesh.triangles.empty()) f.write(reinterpret_cast<const char*>(mesh.triangles.data()), (std::streamsize)(mesh.triangles.size() * sizeof(Triangle)));
        if (!f.good()) { f.close(); std::remove(tmp.c_str()); return; }
    }
    if (std::rename(tmp.c_str(), path.c_str()) != 0) std::remove(tmp.c_str());
}

// cpp:528-586: S * (Rz * Ry * Rx) * T (Q11); trig through double libm of a float radian (Q27)
Raytracer::Matrix Raytracer::ComputeModelMatrix(const Transformation& tr) {
    Matrix S; identity(S);
    S.m[0][0] = tr.scale.x; S.m[1][1] = tr.scale.y; S.m[2][2] = tr.scale.z; S.m[3][3] = 1.0f;
    const float rx = toRadian(tr.rotation.x), ry = toRadian(tr.rotation.y), rz = toRadian(tr.rotation.z);
    Matrix RX; identity(RX);
    // the reference's unqualified cos()/sin() bind to the C library's double functions (Q27)
    auto c = [](float a) { return (float)::cos((double)a); };
    auto s = [](float a) { return (float)::sin((double)a); };
    auto ns = [](float a) { return (float)(-::sin((double)a)); };
    RX.m[1][1] = c(rx); RX.m[1][2] = ns(rx); RX.m[2][1] = s(rx); RX.m[2][2] = c(rx);
    Matrix RY; identity(RY);
    RY.m[0][0] = c(ry); RY.m[0][2] = s(ry); RY.m[2][0] = ns(ry); RY.m[2][2] = c(ry);
    Matrix RZ; identity(RZ);
    RZ.m[0][0] = c(rz); RZ.m[0][1] = ns(rz); RZ.m[1][0] = s(rz); RZ.m[1][1] = c(rz);
    const Matrix R = mul(mul(RZ, RY), RX);                       // cpp:570
    Matrix T; identity(T);
    T.m[0][3] = tr.translation.x; T.m[1][3] = tr.translation.y; T.m[2][3] = tr.translation.z;
    return mul(mul(S, R), T);                                    // cpp:584
}

// cpp:589-643
int Raytracer::LoadMesh(const std::string meshName) {
    if (!mScene) return RT_FAILURE;
    if (mScene->meshMap.find(meshName) != mScene->meshMap.end()) {
        if (!mQuiet) std::cout << "Mesh map already contains " << meshName << ". Skipped loading" << std::endl;
        return RT_SUCCESS;
    }
    std::string text;
    if (!readFile(mAssetsPath + meshName + ".json", text)) {
        std::cout << "File with name " << mAssetsPath << meshName << ".json" << " could not be found";
        return RT_FAILURE;
    }
    std::string cachePath;
    if (!mMeshCacheDir.empty()) {
        cachePath = mMeshCacheDir + (mMeshCacheDir.back() == '/' ? "" : "/") + meshName + ".rt580mesh";
        Mesh cached;
        if (LoadMeshFromCache(cachePath, text, cached)) {        // the cache was written from exactly these JSON bytes
            mScene->meshMap[meshName] = std::move(cached);
            mMeshCacheHits++;
            return RT_SUCCESS;
        }
    }
    jsonmin::ValuePtr doc = jsonmin::parse(text);                // throws like the reference's `file >> jsonData`
    const jsonmin::Value& data = doc->at("data");
    Mesh mesh;
    const std::string shapeType = data.at(0).at("type").as_string();   // cpp:606: data[0] decides (Q26)
    if (shapeType == "polygon") mesh.triangles.reserve(data.size());
    for (size_t it = 0; it < data.size(); it++) {
        const jsonmin::Value& item = data.at(it);
        if (shapeType == "polygon") {
            mesh.type = Mesh::RT_POLYGON;
            Triangle tri;
            for (int i = 0; i < 3; ++i) {
                const jsonmin::Value& v = item.at("v" + std::to_string(i));
                tri.pos[i] = vec3Of(v.at("v"));
                tri.nrm[i] = vec3Of(v.at("n"));
                (void)v.at("t").at(0).as_float(); (void)v.at("t").at(1).as_float();   // cpp:616 requires "t"
            }
            mesh.triangles.push_back(tri);
        } else if (shapeType == "sphere") {
            mesh.type = Mesh::RT_SPHERE;
            mesh.radius = item.at("radius").as_float();
        }
    }
    if (shapeType != "polygon" && shapeType != "sphere") {
        // the reference leaves Mesh::type uninitialised here (Q26); refuse instead of guessing
        std::cout << "Mesh " << meshName << " has unsupported type " << shapeType << "\n";
        return RT_FAILURE;
    }
    if (!cachePath.empty()) StoreMeshInCache(cachePath, text, mesh);
    mScene->meshMap[meshName] = std::move(mesh);
    return RT_SUCCESS;
}

// cpp:645-779
int Raytracer::LoadSceneJSON(const std::string scenePath) {
    int status = 0;
    std::string text;
    if (!readFile(mAssetsPath + scenePath, text)) {
        std::cerr << "Failed to open JSON file" << " Path: " << mAssetsPath << scenePath << "\n";
        return RT_FAILURE;
    }
    jsonmin::ValuePtr doc;
    try {
        doc = jsonmin::parse(text);
    } catch (const std::exception&) {
        std::cout << "Error parsing JSON" << "\n";
        return RT_FAILURE;
    }
    try {
        delete mScene;
        mScene = new Scene();
        mSceneUploaded = false;
        for (size_t k = 0; k < mPeerUploaded.size(); k++) mPeerUploaded[k] = 0;
        const jsonmin::Value& scene = doc->at("scene");
        if (scene.contains("shapes")) {
            const jsonmin::Value& shapes = scene.at("shapes");
            for (size_t si = 0; si < shapes.size(); si++) {
                const jsonmin::Value& sv = shapes.at(si);
                Shape shape;
                shape.id = sv.at("id").as_string();
                shape.geometryId = sv.at("geometry").as_string();
                if (sv.contains("notes")) shape.notes = sv.at("notes").as_string();
                const jsonmin::Value& mat = sv.at("material");
                shape.material.surfaceColor = vec3Of(mat.at("Cs"));
                shape.material.Ka = mat.at("Ka").as_float();
                shape.material.Kd = mat.at("Kd").as_float();
                shape.material.Ks = mat.at("Ks").as_float();
                shape.material.Kt = mat.at("Kt").as_float();
                shape.material.specularExponet = mat.at("n").as_float();      // cpp:685 (Q5)
                const jsonmin::Value& trs = sv.at("transforms");
                for (size_t ti = 0; ti < trs.size(); ti++) {                   // cpp:688-716: last value wins
                    const jsonmin::Value& te = trs.at(ti);
                    if (te.contains("Rx")) shape.transforms.rotation.x = te.at("Rx").as_float();
                    if (te.contains("Ry")) shape.transforms.rotation.y = te.at("Ry").as_float();
                    if (te.contains("Rz")) shape.transforms.rotation.z = te.at("Rz").as_float();
                    if (te.contains("S") && te.at("S").is_array()) shape.transforms.scale = vec3Of(te.at("S"));
                    if (te.contains("T") && te.at("T").is_array()) shape.transforms.translation = vec3Of(te.at("T"));
                }
                mScene->shapes.push_back(shape);
                status |= LoadMesh(shape.geometryId);                          // cpp:719
            }
        }
        if (scene.contains("camera")) {                                        // cpp:724-741
            const jsonmin::Value& cam = scene.at("camera");
            mScene->camera.from = vec3Of(cam.at("from"));
            mScene->camera.to = vec3Of(cam.at("to"));
            for (int i = 0; i < 6; i++) mScene->camera.bounds[i] = cam.at("bounds").at(i).as_float();
            mScene->camera.xRes = cam.at("resolution").at(0).as_int();
            mScene->camera.yRes = cam.at("resolution").at(1).as_int();
        }
        if (scene.contains("lights")) {                                        // cpp:744-771
            const jsonmin::Value& lights = scene.at("lights");
            for (size_t li = 0; li < lights.size(); li++) {
                const jsonmin::Value& lv = lights.at(li);
                Light light;
                light.color = vec3Of(lv.at("color"));
                light.intensity = lv.at("intensity").as_float();
                const std::string typeStr = lv.at("type").as_string();
                if (typeStr == "directional") {
                    light.direction = normalized(sub(vec3Of(lv.at("to")), vec3Of(lv.at("from"))));   // cpp:757-758
                    light.lightType = Light::Directional;
                } else if (typeStr == "ambient") {
                    light.lightType = Light::Ambient;
                } else if (typeStr == "point") {
                    light.lightType = Light::Point;
                    light.position = vec3Of(lv.at("position"));
                } else {
                    // the reference leaves lightType uninitialised; refuse instead
                    throw std::runtime_error("unknown light type '" + typeStr + "'");
                }
                mScene->lights.push_back(light);
            }
        }
        if (!mQuiet) std::cout << "Scene parsing completed!\n";
        mSceneStatus = status;
        if (status == RT_SUCCESS) status |= FlattenScene();
        return status;
    } catch (const std::exception& e) {
        std::cout << "Error parsing JSON " << e.what() << "\n";
        mSceneStatus = RT_FAILURE;
        return RT_FAILURE;
    }
}

// Load-time replacement of the per-ray work at cpp:477-480 and cpp:353-355 (pure => bit-exact):
// world-space vertices, sphere centres (translation column, radius unscaled, Q12), materials
// and lights as SoA float4 records in primitive order (shape order, triangle order).
int Raytracer::FlattenScene() {
    mTriV0.clear(); mTriV1.clear(); mTriV2.clear(); mTriN0.clear(); mTriN1.clear(); mTriN2.clear();
    mSphere.clear(); mMaterials.clear(); mLightF.clear();
    mTriPrim.clear(); mTriMaterial.clear(); mSphPrim.clear(); mSphMaterial.clear(); mLightType.clear();
    size_t nt = 0;
    for (const Shape& sh : mScene->shapes) {
        auto it = mScene->meshMap.find(sh.geometryId);
        if (it == mScene->meshMap.end()) return RT_FAILURE;
        if (it->second.type == Mesh::RT_POLYGON) nt += it->second.triangles.size();
    }
    for (auto* v : { &mTriV0, &mTriV1, &mTriV2, &mTriN0, &mTriN1, &mTriN2 }) v->reserve(nt * 4);
    mTriPrim.reserve(nt); mTriMaterial.reserve(nt);
    int64_t prim = 0;
    int32_t shapeIdx = 0;
    auto push4 = [](Rt580HostVector<float>& v, V3 p) { v.push_back(p.x); v.push_back(p.y); v.push_back(p.z); v.push_back(0.0f); };
    for (const Shape& sh : mScene->shapes) {
        const Mesh& mesh = mScene->meshMap.find(sh.geometryId)->second;
        const Material& m = sh.material;
        const float mrow[8] = { m.surfaceColor.x, m.surfaceColor.y, m.surfaceColor.z, m.Ka, m.Kd, m.Ks, m.Kt, m.specularExponet };
        mMaterials.insert(mMaterials.end(), mrow, mrow + 8);
        const Matrix M = ComputeModelMatrix(sh.transforms);                     // cpp:480
        if (mesh.type == Mesh::RT_POLYGON) {
            for (const Triangle& t : mesh.triangles) {
                push4(mTriV0, transformPoint(M, t.pos[0]));                     // cpp:353
                push4(mTriV1, transformPoint(M, t.pos[1]));                     // cpp:354
                push4(mTriV2, transformPoint(M, t.pos[2]));                     // cpp:355
                push4(mTriN0, t.nrm[0]); push4(mTriN1, t.nrm[1]); push4(mTriN2, t.nrm[2]);   // object space (Q10)
                mTriPrim.push_back((int32_t)prim++);
                mTriMaterial.push_back(shapeIdx);
            }
        } else {
            mSphere.push_back(M.m[0][3]); mSphere.push_back(M.m[1][3]); mSphere.push_back(M.m[2][3]);   // h:212-214
            mSphere.push_back(mesh.radius);                                     // cpp:423, unscaled
            mSphPrim.push_back((int32_t)prim++);
            mSphMaterial.push_back(shapeIdx);
        }
        shapeIdx++;
    }
    mNumPrims = prim;
    for (const Light& l : mScene->lights) {
        mLightType.push_back(l.lightType == Light::Directional ? RT580_LIGHT_DIRECTIONAL
                             : l.lightType == Light::Point     ? RT580_LIGHT_POINT : RT580_LIGHT_AMBIENT);
        const float row[10] = { l.color.x, l.color.y, l.color.z, l.intensity, l.position.x, l.position.y, l.position.z,
                                l.direction.x, l.direction.y, l.direction.z };
        mLightF.insert(mLightF.end(), row, row + 10);
    }
    mSceneUploaded = false;
    for (size_t k = 0; k < mPeerUploaded.size(); k++) mPeerUploaded[k] = 0;
    return RT_SUCCESS;
}

int Raytracer::GetInstancedScene(rt580_instanced_scene* out) {
    if (!out || !mScene) return RT_INVALID_ARG;
    memset(out, 0, sizeof *out);
    mInstMeshFirst.assign(1, 0); mInstMeshTris.clear(); mInstMatrix.clear(); mInstRadius.clear(); mInstShapeMesh.clear();
    std::map<std::string, int32_t> meshIndex;
    for (const Shape& sh : mScene->shapes) {
        auto it = mScene->meshMap.find(sh.geometryId);
        if (it == mScene->meshMap.end()) return RT_FAILURE;
        const Mesh& mesh = it->second;
        int32_t mi = -1;
        if (mesh.type == Mesh::RT_POLYGON) {
            auto f = meshIndex.find(sh.geometryId);
            if (f == meshIndex.end()) {
                mi = (int32_t)meshIndex.size();
                meshIndex[sh.geometryId] = mi;
                static_assert(sizeof(Triangle) == 18 * sizeof(float), "Triangle is 18 packed floats");
                const float* p = reinterpret_cast<const float*>(mesh.triangles.data());
                mInstMeshTris.insert(mInstMeshTris.end(), p, p + mesh.triangles.size() * 18);
                mInstMeshFirst.push_back((int64_t)(mInstMeshTris.size() / 18));
            } else mi = f->second;
        }
        mInstShapeMesh.push_back(mi);
        const Matrix M = ComputeModelMatrix(sh.transforms);                     // cpp:480
        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) mInstMatrix.push_back(M.m[r][c]);
        mInstRadius.push_back(mesh.radius);
    }
    out->n_meshes = (int32_t)meshIndex.size();
    out->mesh_first = mInstMeshFirst.data(); out->mesh_tris = mInstMeshTris.data();
    out->n_shapes = (int32_t)mInstShapeMesh.size();
    out->shape_mesh = mInstShapeMesh.data(); out->shape_matrix = mInstMatrix.data(); out->shape_radius = mInstRadius.data();
    out->materials = mMaterials.data();
    out->n_lights = (int32_t)mLightType.size();
    out->light_type = mLightType.data(); out->light_f = mLightF.data();
    out->origin_hint[0] = mScene->camera.from.x; out->origin_hint[1] = mScene->camera.from.y; out->origin_hint[2] = mScene->camera.from.z;
    return RT_SUCCESS;
}

int Raytracer::GetFlatScene(rt580_flat_scene* out) const {
    if (!out || !mScene) return RT_INVALID_ARG;
    memset(out, 0, sizeof *out);
    out->n_prims = mNumPrims;
    out->n_tris = (int64_t)mTriPrim.size();
    out->tri_v0 = mTriV0.data(); out->tri_v1 = mTriV1.data(); out->tri_v2 = mTriV2.data();
    out->tri_n0 = mTriN0.data(); out->tri_n1 = mTriN1.data(); out->tri_n2 = mTriN2.data();
    out->tri_prim = mTriPrim.data(); out->tri_material = mTriMaterial.data();
    out->n_spheres = (int64_t)mSphPrim.size();
    out->sph_center_r = mSphere.data(); out->sph_prim = mSphPrim.data(); out->sph_material = mSphMaterial.data();
    out->n_materials = (int32_t)(mMaterials.size() / 8);
    out->materials = mMaterials.data();
    out->n_lights = (int32_t)mLightType.size();
    out->light_type = mLightType.data(); out->light_f = mLightF.data();
    out->origin_hint[0] = mScene->camera.from.x; out->origin_hint[1] = mScene->camera.from.y; out->origin_hint[2] = mScene->camera.from.z;
    return RT_SUCCESS;
}

// cpp:895-915 + cpp:861-870 (+ the per-pixel Matrix::Inverse of cpp:849-850, hoisted: pure)
int Raytracer::InitializeRenderer() {
    if (!mScene) return RT_FAILURE;
    Camera& cam = mScene->camera;
    const V3 n = normalized(sub(cam.from, cam.to));
    const V3 worldUp = mk(0, 1, 0);
    const V3 u = normalized(cross3(worldUp, n));
    const V3 v = normalized(cross3(n, u));
    const V3 r = cam.from;
    Matrix& view = cam.viewMatrix;
    view.m[0][0] = u.x; view.m[0][1] = u.y; view.m[0][2] = u.z; view.m[0][3] = -dot3(r, u);
    view.m[1][0] = v.x; view.m[1][1] = v.y; view.m[1][2] = v.z; view.m[1][3] = -dot3(r, v);
    view.m[2][0] = n.x; view.m[2][1] = n.y; view.m[2][2] = n.z; view.m[2][3] = -dot3(r, n);
    view.m[3][0] = 0; view.m[3][1] = 0; view.m[3][2] = 0; view.m[3][3] = 1;
    // CalculateProjectionMatrix (cpp:880-893) is never read by the render path (Q23): skipped.
    Matrix inv;
    mViewOk = inverse4(view, inv) == RT_SUCCESS;
    if (!mViewOk) {
        std::cerr << "Failed to compute the inverse of the view matrix.\n";           // cpp:856
        for (float& f : mInvView) f = 0.0f;     // the reference then traces zero-direction rays
    } else {
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) mInvView[3 * i + j] = inv.m[i][j];
    }
    return RT_SUCCESS;                           // cpp:912-914: both helpers always succeed
}

int Raytracer::GetRenderParams(rt580_render_params* out) {
    if (!out || !mScene) return RT_INVALID_ARG;
    InitializeRenderer();
    memset(out, 0, sizeof *out);
    out->width = mWidth; out->height = mHeight; out->fov_degrees = mFov;
    out->camera_from[0] = mScene->camera.from.x; out->camera_from[1] = mScene->camera.from.y; out->camera_from[2] = mScene->camera.from.z;
    memcpy(out->inv_view3x3, mInvView, sizeof mInvView);
    out->depth = mDepth; out->ao_spp = mAoSpp; out->rng_mode = mRngMode; out->traversal = mTraversal;
    out->row_first = 0; out->row_step = 1; out->n_rows = 0;
    out->farfield = mFarField;
    return RT_SUCCESS;
}

int Raytracer::EnsureContext() {
    if (!mCtx) {
        if (rt580_create(mDevice, &mCtx) != RT580_SUCCESS) {
            std::cerr << "Raytracer: " << rt580_last_error() << "\n";
            mCtx = nullptr;
            return RT_FAILURE;
        }
    }
    if (!mSceneUploaded) {
        rt580_flat_scene fs;
        rt580_instanced_scene is;
        if (GetFlatScene(&fs) != RT_SUCCESS) return RT_FAILURE;
        if (mDeviceFlatten && GetInstancedScene(&is) != RT_SUCCESS) return RT_FAILURE;
        if ((mDeviceFlatten ? rt580_upload_instanced_scene(mCtx, &is) : rt580_upload_scene(mCtx, &fs)) != RT580_SUCCESS) {
            std::cerr << "Raytracer: " << rt580_last_error() << "\n";
            return RT_FAILURE;
        }
        mSceneUploaded = true;
    }
    return RT_SUCCESS;
}

int Raytracer::RenderToFrameBuffer() {
    if (!mScene || mSceneStatus != RT_SUCCESS) {
        std::cerr << "Raytracer: no scene loaded\n";
        return RT_FAILURE;
    }
    if (mWidth <= 0 || mHeight <= 0) return RT_INVALID_ARG;
    rt580_render_params rp;
    if (GetRenderParams(&rp) != RT_SUCCESS) return RT_FAILURE;      // InitializeRenderer(), cpp:917
    if (EnsureContext() != RT_SUCCESS) return RT_FAILURE;
    if (mGpus > 1) return RenderMultiGpu(rp);
    static_assert(sizeof(Pixel) == 6, "Pixel must be 3 packed shorts (h:373-374)");
    const int st = rt580_render(mCtx, &rp, reinterpret_cast<int16_t*>(mFrameBuffer.data()), &mStats);   // cpp:921-932
    if (st != RT580_SUCCESS) std::cerr << "Raytracer: " << rt580_last_error() << "\n";
    return st;
}

// The frame on several GPUs of this process: rows interleaved (row y on GPU y % n), the structure pass of every GPU first, then
// the exchange - the hit-node counts per row, prefix-summed in scanline order, give every row its place in the reference's
// single random stream (h:592) - then the occlusion + resolve passes, each GPU's rows into the frame buffer.
int Raytracer::RenderMultiGpu(const rt580_render_params& rp) {
    const int n = mGpus;
    mPeers.resize((size_t)n - 1, nullptr);
    mPeerUploaded.resize((size_t)n - 1, 0);
    rt580_flat_scene fs;
    rt580_instanced_scene is;
    if (GetFlatScene(&fs) != RT_SUCCESS) return RT_FAILURE;
    if (mDeviceFlatten && GetInstancedScene(&is) != RT_SUCCESS) return RT_FAILURE;
    std::vector<rt580_context*> ctx((size_t)n);
    ctx[0] = mCtx;
    std::vector<int> status((size_t)n, RT580_SUCCESS);
    std::vector<std::string> errs((size_t)n);
    auto on_all = [&](auto fn) {
        std::vector<std::thread> th;
        for (int g = 0; g < n; g++) th.emplace_back([&, g] { status[g] = fn(g); if (status[g] != RT580_SUCCESS) errs[g] = rt580_last_error(); });
        for (auto& t : th) t.join();
        for (int g = 0; g < n; g++) if (status[g] != RT580_SUCCESS) { std::cerr << "Raytracer: GPU " << g << ": " << errs[g] << "\n"; return status[g]; }
        return (int)RT580_SUCCESS;
    };
    int st = on_all([&](int g) -> int {
        if (g == 0) return RT580_SUCCESS;
        if (!mPeers[g - 1]) { const int s2 = rt580_create(mDevice + g, &mPeers[g - 1]); if (s2 != RT580_SUCCESS) return s2; }
        ctx[g] = mPeers[g - 1];
        if (!mPeerUploaded[g - 1]) { const int s2 = mDeviceFlatten ? rt580_upload_instanced_scene(ctx[g], &is) : rt580_upload_scene(ctx[g], &fs); if (s2 != RT580_SUCCESS) return s2; mPeerUploaded[g - 1] = 1; }
        return RT580_SUCCESS;
    });
    if (st != RT580_SUCCESS) return st;
    std::vector<rt580_render_params> ps((size_t)n, rp);
    std::vector<std::vector<uint64_t>> counts((size_t)n);
    for (int g = 0; g < n; g++) {
        const int rows = g < mHeight ? (mHeight - g + n - 1) / n : 0;
        ps[g].row_first = rows ? g : 0; ps[g].row_step = n; ps[g].n_rows = rows ? rows : -1;
        counts[g].assign((size_t)rows + 1, 0);
    }
    st = on_all([&](int g) { return rt580_render_begin(ctx[g], &ps[g], counts[g].data()); });
    if (st != RT580_SUCCESS) return st;
    uint64_t run = 0;                                    // exclusive prefix over the rows in scanline order (cpp:921-922)
    for (int y = 0; y < mHeight; y++) { uint64_t& c = counts[y % n][y / n]; const uint64_t v = c; c = run; run += v; }
    std::vector<Rt580HostVector<int16_t>> bands((size_t)n);
    std::vector<rt580_stats> stats((size_t)n);
    for (int g = 0; g < n; g++) bands[g].resize((size_t)(ps[g].n_rows > 0 ? ps[g].n_rows : 0) * mWidth * 3 + 1);
    st = on_all([&](int g) { return rt580_render_finish(ctx[g], counts[g].data(), bands[g].data(), 0, &stats[g]); });
    if (st != RT580_SUCCESS) return st;
    int16_t* fb = reinterpret_cast<int16_t*>(mFrameBuffer.data());
    for (int y = 0; y < mHeight; y++)
        memcpy(fb + (size_t)y * mWidth * 3, bands[y % n].data() + (size_t)(y / n) * mWidth * 3, (size_t)mWidth * 6);
    mStats = stats[0];
    for (int g = 1; g < n; g++) {                        // rays add up, times are the slowest GPU's
        mStats.rays_primary += stats[g].rays_primary; mStats.rays_secondary += stats[g].rays_secondary;
        mStats.rays_shadow += stats[g].rays_shadow; mStats.rays_ao += stats[g].rays_ao;
        mStats.hit_nodes += stats[g].hit_nodes; mStats.ao_calls += stats[g].ao_calls;
        mStats.far_scans += stats[g].far_scans; mStats.linear_fallbacks += stats[g].linear_fallbacks;
        mStats.ao_rays_traversed += stats[g].ao_rays_traversed; mStats.shadow_rays_traversed += stats[g].shadow_rays_traversed;
        mStats.kernel_launches += stats[g].kernel_launches;
        if (stats[g].ms_total > mStats.ms_total) {
            mStats.ms_total = stats[g].ms_total; mStats.ms_structure = stats[g].ms_structure; mStats.ms_order = stats[g].ms_order;
            mStats.ms_ao = stats[g].ms_ao; mStats.ms_resolve = stats[g].ms_resolve; mStats.ms_ao_kernel = stats[g].ms_ao_kernel;
        }
    }
    return RT_SUCCESS;
}

// cpp:916-935
int Raytracer::Render(const std::string outputName) {
    const int st = RenderToFrameBuffer();
    if (st != RT_SUCCESS) return st;
    // cpp:934 FlushFrameBufferToPPM: the gamma table is applied on the device and the PPM body comes back
    // as 3 bytes per pixel (rt580_frame_rgb8); the host restatement below is the fallback and the
    // public method (it writes whatever the frame buffer holds, like the reference's)
    if (mCtx && mGpus == 1 && mWidth > 0 && mHeight > 0) {
        unsigned char lut[256];
        for (int c = 0; c < 256; c++) lut[c] = static_cast<unsigned char>(std::pow(c / 255.0f, 1.0f / 2.2f) * 255.0f);   // cpp:816-818
        Rt580HostVector<unsigned char> body((size_t)mWidth * mHeight * 3);
        if (rt580_frame_rgb8(mCtx, lut, body.data(), 0) == RT580_SUCCESS) {
            std::ofstream outfile(outputName, std::ios::binary);
            if (!outfile.is_open()) {
                std::cerr << "Failed to create output file: " << outputName << std::endl;
                return RT_FAILURE;
            }
            outfile << "P6\n" << mWidth << " " << mHeight << "\n255\n";
            outfile.write(reinterpret_cast<const char*>(body.data()), (std::streamsize)body.size());
            outfile.close();
            return RT_SUCCESS;
        }
    }
    return FlushFrameBufferToPPM(outputName);
}

// cpp:796-830: gamma 1/2.2 through libm powf, truncation to 8 bit (Q24)
int Raytracer::FlushFrameBufferToPPM(std::string outputName) {
    if (mFrameBuffer.empty()) {
        std::cerr << "Display or frame buffer is null." << std::endl;
        return RT_FAILURE;
    }
    std::ofstream outfile(outputName, std::ios::binary);
    if (!outfile.is_open()) {
        std::cerr << "Failed to create output file: " << outputName << std::endl;
        return RT_FAILURE;
    }
    outfile << "P6\n" << mWidth << " " << mHeight << "\n255\n";
    // a 256-entry table of the same expression is exact for the clamped values the path emits
    unsigned char lut[256];
    for (int c = 0; c < 256; c++) lut[c] = static_cast<unsigned char>(std::pow(c / 255.0f, 1.0f / 2.2f) * 255.0f);
    std::vector<unsigned char> row((size_t)mWidth * 3);
    for (int y = 0; y < mHeight; y++) {
        for (int x = 0; x < mWidth; x++) {
            const Pixel& p = mFrameBuffer[(size_t)y * mWidth + x];
            const short ch[3] = { p.r, p.g, p.b };
            for (int k = 0; k < 3; k++) {
                const short c = ch[k];
                row[3 * (size_t)x + k] = (c >= 0 && c <= 255) ? lut[c]
                                        : static_cast<unsigned char>(std::pow(c / 255.0f, 1.0f / 2.2f) * 255.0f);
            }
        }
        outfile.write(reinterpret_cast<const char*>(row.data()), (std::streamsize)row.size());
    }
    outfile.close();
    return RT_SUCCESS;
}
