// Host-side Raytracer class: the reference's public API (Raytracer.h:557-588) kept as is,
// so main() at Raytracer.cpp:944-953 compiles against it unchanged:
//
//     Raytracer rt(500, 500);
//     rt.LoadSceneJSON("simpleSphereScene.json");
//     rt.Render("output.ppm");
//
// LoadSceneJSON / LoadMesh (cpp:645-779, 589-643), InitializeRenderer (cpp:895-915) and
// FlushFrameBufferToPPM (cpp:796-830) stay host C++ (this repo's own code); Render
// (cpp:916-935) dispatches the per-pixel work through the C ABI of include/rt580.h to the
// CUDA library.  New, and not in the reference: FlattenScene(), run at the end of
// LoadSceneJSON, which performs once per scene what the reference redoes per ray
// (ComputeModelMatrix cpp:480, TransformPoint cpp:353-355).
//
// Error behaviour matches the reference: int status RT_SUCCESS / RT_FAILURE / RT_INVALID_ARG
// (h:8-10), OR-accumulated, diagnostics on cout / cerr, no exception escapes.
#pragma once
#include <cstdint>
#include <map>
#include <new>
#include <string>
#include <vector>
#include "../../include/rt580.h"

// std::vector over rt580_host_alloc (page-locked when a GPU is present): the flattened scene and the
// frame buffer cross PCIe at full speed
template <class T> struct Rt580HostAlloc {
    using value_type = T;
    Rt580HostAlloc() = default;
    template <class U> Rt580HostAlloc(const Rt580HostAlloc<U>&) {}
    T* allocate(size_t n) { void* p = rt580_host_alloc((uint64_t)n * sizeof(T)); if (!p) throw std::bad_alloc(); return static_cast<T*>(p); }
    void deallocate(T* p, size_t) { rt580_host_free(p); }
    template <class U> bool operator==(const Rt580HostAlloc<U>&) const { return true; }
    template <class U> bool operator!=(const Rt580HostAlloc<U>&) const { return false; }
};
template <class T> using Rt580HostVector = std::vector<T, Rt580HostAlloc<T>>;

#define RT_SUCCESS      0
#define RT_FAILURE      1
#define RT_INVALID_ARG  2

class Raytracer {
public:
    struct Vector3 { float x = 0.0f, y = 0.0f, z = 0.0f; };
    struct Matrix { float m[4][4]; };
    struct Pixel { short r = 0, g = 0, b = 0; };                       // h:373-374
    struct Material {                                                  // h:442-463
        Vector3 surfaceColor{ 1.0f, 1.0f, 1.0f };
        float Ka = 0.5f, Kd = 0.75f, Ks = 0.95f, Kt = 0.95f;
        float refractiveIndex = 2.5f;                                  // never loaded (Q5)
        float specularExponet = 32.0f;
    };
    struct Transformation {                                            // h:532-537
        Vector3 scale{ 1.0f, 1.0f, 1.0f };
        Vector3 rotation;      // degrees about x, y, z
        Vector3 translation;
    };
    struct Triangle { Vector3 pos[3]; Vector3 nrm[3]; };               // h:436-469 (texture coords unused by the path)
    struct Mesh { enum Type { RT_POLYGON, RT_SPHERE } type = RT_POLYGON; std::vector<Triangle> triangles; float radius = 0.0f; };
    struct Shape { std::string id, geometryId, notes; Material material; Transformation transforms; };
    struct Light {                                                     // h:519-530
        enum Type { Directional, Point, Ambient } lightType = Ambient;
        Vector3 color; float intensity = 0.0f; Vector3 position; Vector3 direction;
    };
    struct Camera { Matrix viewMatrix; Vector3 from, to; float bounds[6] = { 0 }; int xRes = 0, yRes = 0; };
    struct Scene {
        std::vector<Shape> shapes;
        Camera camera;
        std::map<std::string, Mesh> meshMap;
        std::vector<Light> lights;
    };

    Raytracer(int width, int height);                                  // h:588, cpp:781-788
    ~Raytracer();
    Raytracer(const Raytracer&) = delete;
    Raytracer& operator=(const Raytracer&) = delete;

    int LoadSceneJSON(const std::string scenePath);                    // h:572
    int LoadMesh(const std::string meshName);                          // h:571
    int Render(const std::string outputName);                          // h:586
    int FlushFrameBufferToPPM(std::string outputName);                 // h:573
    Matrix ComputeModelMatrix(const Transformation& transform);        // h:574

    // ---- not in the reference: what its hard-coded constants become --------------------------
    void SetAssetsPath(const std::string& dir) { mAssetsPath = dir; }  // ASSETS_PATH h:15
    void SetBounces(int depth) { mDepth = depth; }                     // default argument at h:563
    void SetAmbientOcclusionSamples(int spp) { mAoSpp = spp; }         // literal at cpp:317
    void SetRngMode(int mode) { mRngMode = mode; }
    void SetTraversal(int t) { mTraversal = t; }
    void SetDevice(int device) { mDevice = device; }
    void SetFarField(int mode) { mFarField = mode; }
    void SetQuiet(bool q) { mQuiet = q; }                              // silence the loader's cout chatter (cpp:592, 772)
    // Rows of the frame interleaved over the first n GPUs of this process (row y on GPU y % n), one context each; the per-row
    // hit-node counts that position every row in the reference's single random stream (h:592) are exchanged on the host.
    void SetGpus(int n) { mGpus = n < 1 ? 1 : n; }
    // Binary cache of parsed meshes: "<dir>/<mesh>.rt580mesh" next to / instead of re-parsing "<mesh>.json"
    // (teapot.json: 472 KB of text for 1024 triangles).  Keyed by the JSON's size and a hash of its bytes: the JSON stays the
    // source of truth, a stale or damaged cache file is ignored and rewritten.  Empty = no cache (the default).
    void SetMeshCacheDir(const std::string& dir) { mMeshCacheDir = dir; }
    // FlattenScene's TransformPoint calls (cpp:353-355) on the device: the context receives the meshes in object
    // space and one model matrix per shape instead of the flattened arrays.  Same bytes on the device, same frame.
    void SetDeviceFlatten(bool on) { if (on != mDeviceFlatten) { mDeviceFlatten = on; mSceneUploaded = false; for (auto& u : mPeerUploaded) u = 0; } }
    int  GetInstancedScene(rt580_instanced_scene* out);
    int  MeshCacheHits() const { return mMeshCacheHits; }
    rt580_context* Context() const { return mCtx; }
    int  RenderToFrameBuffer();                                        // Render without the PPM
    const Pixel* FrameBuffer() const { return mFrameBuffer.data(); }
    int Width() const { return mWidth; }
    int Height() const { return mHeight; }
    const rt580_stats& Stats() const { return mStats; }
    const Scene* GetScene() const { return mScene; }
    int GetFlatScene(rt580_flat_scene* out) const;
    int GetRenderParams(rt580_render_params* out);

private:
    int InitializeRenderer();                                          // h:604
    int FlattenScene();
    int EnsureContext();
    int RenderMultiGpu(const rt580_render_params& rp);
    bool LoadMeshFromCache(const std::string& path, const std::string& jsonText, Mesh& mesh);
    void StoreMeshInCache(const std::string& path, const std::string& jsonText, const Mesh& mesh);

    std::string mAssetsPath = "Assets/";
    int mWidth, mHeight;
    float mFov = 60.0f;                                                // cpp:786
    Rt580HostVector<Pixel> mFrameBuffer;                               // Display::frameBuffer h:423
    Scene* mScene = nullptr;
    int mSceneStatus = RT_SUCCESS;
    int mDepth = 4, mAoSpp = 128, mRngMode = RT580_RNG_REFERENCE_LCG, mTraversal = RT580_TRAVERSAL_AUTO, mDevice = 0,
        mFarField = RT580_FARFIELD_EXACT;
    float mInvView[9] = { 0 };
    bool mViewOk = false;
    bool mQuiet = false;

    // flattened scene (host SoA, float4 records) handed to rt580_upload_scene
    Rt580HostVector<float> mTriV0, mTriV1, mTriV2, mTriN0, mTriN1, mTriN2;
    std::vector<float> mSphere, mMaterials, mLightF;
    Rt580HostVector<int32_t> mTriPrim, mTriMaterial;
    std::vector<int32_t> mSphPrim, mSphMaterial, mLightType;
    int64_t mNumPrims = 0;

    rt580_context* mCtx = nullptr;
    std::vector<rt580_context*> mPeers;    // contexts on GPUs 1 .. mGpus-1 (SetGpus)
    std::vector<char> mPeerUploaded;      // char, not bool: written from one thread per GPU
    int mGpus = 1;
    bool mDeviceFlatten = false;
    // the un-flattened scene for rt580_upload_instanced_scene (built on demand by GetInstancedScene)
    std::vector<int64_t> mInstMeshFirst; std::vector<float> mInstMeshTris, mInstMatrix, mInstRadius; std::vector<int32_t> mInstShapeMesh;
    std::string mMeshCacheDir;
    int mMeshCacheHits = 0;
    bool mSceneUploaded = false;
    rt580_stats mStats{};
};
