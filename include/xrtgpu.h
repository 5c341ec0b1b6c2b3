/*
 * xrtgpu.h — C ABI of libxrtgpu.so, the B200 (sm_100a) wavefront path tracer that sits
 * behind the xRayTracer `Renderer` plug-in point.
 *
 * Reference interface each entry point replaces (paths relative to /root/reference/Src):
 *   xrtg_scene_create[2]  <- Scene::loadObj/addObj/addAreaLight/addDeltaLight + the empty hook
 *                            Scene::build()                      scene.h:13-30, scene.cpp:46-170
 *   xrtg_render[_device]  <- Renderer::render / NormalRenderer::doRender / ParallelRenderer::render
 *                                                               renderer.h:8-20, renderer.cpp:8-99
 *   xrtg_trace_primary    <- PinholeCamera::sampleRay + Scene::intersect (parity hook)
 *                                                               camera.h:49-60, scene.cpp:190-200
 *   xrtg_trace_rays       <- Scene::intersect / Scene::occluded  scene.cpp:190-211
 *   xrtg_image_to_u8      <- Image::gammaCorrection + writePPM / writeMat quantisation   image.h:80-136
 *   xrtg_scene_create_multi, xrtg_reduce_finalize, xrtg_exchange_buffer, xrtg_ipc_*
 *                         <- the "one renderer, every core" role of ParallelRenderer::render (renderer.cpp:83-99) scaled to
 *                            several GPUs: samples split across devices, `image /= n_samples` (renderer.cpp:98) fused into the
 *                            peer-memory reduction of the per-device sums
 *
 * Conventions
 *   - every function returns 0 on success or a negative xrtg_status; xrtg_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread.
 *   - plain pointers and sizes only; no C++/torch types cross this boundary.
 *   - all arithmetic on the path is fp32 (the reference's Vec3f, geometry.h:263).
 *   - image layout is the reference's: row-major RGB, index = j + W*i (image.h:140-143).
 *   - objects[] are listed in the reference's *iteration* order (std::unordered_map order,
 *     scene.cpp:193) and global primitive ids are assigned by walking objects[] in that order
 *     (mesh: one id per triangle, sphere/box: one id). Ties in t resolve to the lowest id, which
 *     is exactly what the reference's "first strictly smaller t wins" loops produce.
 *   - there is no CPU fallback anywhere behind this ABI: without a CUDA device every compute entry
 *     point fails with XRTG_ERR_NO_DEVICE.
 */
#ifndef XRTGPU_H
#define XRTGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this ABI is exported */
#endif

#define XRTG_ABI_VERSION 2

typedef enum xrtg_status {
    XRTG_OK = 0,
    XRTG_ERR_INVALID = -1,     /* bad argument / malformed scene description            */
    XRTG_ERR_NO_DEVICE = -2,   /* no CUDA device (there is no CPU fallback)              */
    XRTG_ERR_CUDA = -3,        /* a CUDA runtime call failed; message has the error      */
    XRTG_ERR_UNSUPPORTED = -4, /* integrator / scene feature outside the GPU path        */
    XRTG_ERR_OOM = -5
} xrtg_status;

/* ---- geometry PODs --------------------------------------------------------------------- */

/* One reference `Primitive` (primitive.h:6-35): 3 positions + 3 shading normals.
 * Texcoords are not on the path (Lambert ignores them, material.h:39-48). */
typedef struct xrtg_triangle {
    float v0[3], v1[3], v2[3];
    float n0[3], n1[3], n2[3];
} xrtg_triangle;

/* `Sphere` (primitive.h:97-184). */
typedef struct xrtg_sphere {
    float center[3];
    float radius;
} xrtg_sphere;

/* `BoxMesh` (primitive.h:230-273), the bounding proxy of a medium. */
typedef struct xrtg_box {
    float pmin[3], pmax[3];
} xrtg_box;

typedef enum xrtg_object_kind { XRTG_OBJ_MESH = 0, XRTG_OBJ_SPHERE = 1, XRTG_OBJ_BOX = 2 } xrtg_object_kind;

/* `Object` facade (primitive.h:40-95). */
typedef struct xrtg_object {
    int32_t kind;       /* xrtg_object_kind                                              */
    int32_t first;      /* first triangle / sphere / box in the matching array          */
    int32_t count;      /* number of triangles (mesh) or 1                              */
    int32_t material;   /* index into materials[], -1 = nullptr (light proxies, media)   */
    int32_t area_light; /* index into area_lights[], -1 = none                           */
    int32_t medium;     /* index into media[], -1 = none                                 */
    int32_t insert_seq; /* order in which the host Scene inserted it (addObj order)      */
    int32_t _pad;
    const char* name;   /* key in the reference's m_objects map; may be NULL            */
} xrtg_object;

typedef enum xrtg_material_kind { XRTG_MAT_LAMBERT = 0 } xrtg_material_kind;

/* `Lambert` (material.h:28-77). */
typedef struct xrtg_material {
    int32_t kind;
    float albedo[3];
} xrtg_material;

typedef enum xrtg_area_light_kind {
    XRTG_LIGHT_QUAD = 0,     /* light.cpp:49-82  */
    XRTG_LIGHT_TRIANGLE = 1, /* light.cpp:6-47   */
    XRTG_LIGHT_SPHERE = 2    /* light.h:124-198 (cone-sampling branch), light.cpp:84-113 */
} xrtg_area_light_kind;

/* World-space (already multiplied by lightToWorld). Quad/triangle: v0,v1,v2; e1=v1-v0, e2=v2-v0,
 * Ng=e1 x e2 are derived. Sphere: v0 = centre, radius. */
typedef struct xrtg_area_light {
    int32_t kind;
    float v0[3], v1[3], v2[3];
    float radius;
    float Le[3];
} xrtg_area_light;

typedef enum xrtg_delta_light_kind { XRTG_DLIGHT_POINT = 0, XRTG_DLIGHT_DISTANT = 1 } xrtg_delta_light_kind;

/* `PointLight` / `DistantLight` (light.cpp:115-142). radiance = color * intensity. */
typedef struct xrtg_delta_light {
    int32_t kind;
    float pos_or_dir[3]; /* point: world position; distant: normalised travel direction `dir` */
    float radiance[3];
} xrtg_delta_light;

typedef enum xrtg_medium_kind {
    XRTG_MEDIUM_HOMOGENEOUS_MIS = 0,        /* medium.h:148-192 */
    XRTG_MEDIUM_HOMOGENEOUS_ACHROMATIC = 1, /* medium.h:195-229 */
    XRTG_MEDIUM_HOMOGENEOUS_NOMIS = 2,      /* medium.h:232-277 */
    XRTG_MEDIUM_HETEROGENEOUS = 3           /* medium.h:280-387, medium.cpp:5-133 */
} xrtg_medium_kind;

typedef struct xrtg_medium {
    int32_t kind;
    float g;          /* Henyey-Greenstein asymmetry (medium.h:21-68)                    */
    float sigma_a[3]; /* homogeneous: absorption;  heterogeneous: absorptionColor        */
    float sigma_s[3]; /* homogeneous: scattering;  heterogeneous: scatteringColor        */
    float density_mul;
    int32_t grid;     /* heterogeneous: index into grids[], else -1                      */
} xrtg_medium;

/* Dense restatement of `DensityGrid`/`OpenVDBGrid` (grid.h:9-85): fp32 voxels at integer index
 * coordinates [0,n) (x fastest), world = origin + voxel_size * index, trilinear (BoxSampler) lookup,
 * `background` outside. bounds = indexToWorld of the active-voxel bbox [active_min, active_max]. */
typedef struct xrtg_grid {
    int32_t nx, ny, nz;
    const float* data; /* host pointer, nx*ny*nz floats */
    float origin[3];
    float voxel_size;
    float background;
    int32_t active_min[3], active_max[3]; /* inclusive index bbox of non-background voxels */
    float max_density;                    /* evalMinMax max (grid.h:79-84)                 */
} xrtg_grid;

typedef struct xrtg_scene_desc {
    int32_t abi_version; /* XRTG_ABI_VERSION */
    int32_t n_objects, n_triangles, n_spheres, n_boxes;
    int32_t n_materials, n_area_lights, n_delta_lights, n_media, n_grids;
    const xrtg_object* objects; /* reference iteration order */
    const xrtg_triangle* triangles;
    const xrtg_sphere* spheres;
    const xrtg_box* boxes;
    const xrtg_material* materials;
    const xrtg_area_light* area_lights;   /* Scene::getAreaLights() order (scene.cpp:166-170) */
    const xrtg_delta_light* delta_lights; /* Scene::getDeltaLights() order                    */
    const xrtg_medium* media;
    const xrtg_grid* grids;
} xrtg_scene_desc;

/* `PinholeCamera` (camera.h:34-60): c2w row-major with translation in row 3, scale = tan(FOV/2)
 * evaluated on the host, aspect = W/H. */
typedef struct xrtg_camera {
    float c2w[16];
    float scale;
    float aspect;
} xrtg_camera;

typedef enum xrtg_integrator {
    XRTG_INT_NORMAL = 0,   /* NormalIntegrator as shipped: 0.5*(ns+1)       integrator.h:29-36  */
    XRTG_INT_FURNACE = 1,  /* the dead furnace block                        integrator.h:59-66  */
    XRTG_INT_DIRECT = 2,   /* DirectIntegrator                              integrator.h:82-119 */
    XRTG_INT_INDIRECT = 3, /* IndirectIntegrator                            integrator.h:129-186 */
    XRTG_INT_GI = 4,       /* GIIntegrator                                  integrator.h:205-287 */
    XRTG_INT_WHITTED = 5,  /* WhittedIntegrator, Lambert/delta-light branch integrator.h:328-343 */
    XRTG_INT_VOLUME = 6,   /* VolumePathTracing                             integrator.h:409-473 */
    XRTG_INT_VOLUME_NEE = 7 /* VolumePathTracingNEE                         integrator.h:489-631 */
} xrtg_integrator;

enum {
    /* Reproduce the reference's sample stream: per-pixel std::mt19937 seeded j+W*i (renderer.cpp:35-36),
     * libstdc++ uniform_real_distribution<float> mapping, no FMA contraction, one sample per pixel per
     * wave. Slow; exists for parity. Without it: counter-based RNG keyed (seed,pixel,sample,dim). */
    XRTG_FLAG_EXACT = 1u << 0,
    /* Count closest-hit rays, shadow rays, BVH nodes visited, triangles tested, tracking steps. */
    XRTG_FLAG_COUNTERS = 1u << 1,
    /* Closest/any-hit by brute force in primitive order instead of the BVH (parity debugging). */
    XRTG_FLAG_BRUTE_FORCE = 1u << 2,
    /* Do not divide by spp: leave the per-pixel SUM in the output (multi-GPU partial results). */
    XRTG_FLAG_SUM_ONLY = 1u << 3,
    /* Bracket every kernel launch with CUDA events and report per-stage device time in xrtg_stats
     * (extend_ms / shade_ms / connect_ms / other_ms). Cheap: no extra work inside the kernels. */
    XRTG_FLAG_STAGE_TIMES = 1u << 4,
    /* Parity hooks only (xrtg_trace_primary / xrtg_trace_rays): trace with the THROUGHPUT instantiation through the very entry
     * points xrtg_render uses for this scene — k_primary incl. its screen-space scissor, the shared-memory small-scene tracer of
     * k_bounce_small (plane-paired records, hull-pruned occluders), the simple kernels on mid-size scenes, k_trace on the wide
     * BVH with plane-equation triangle records on deep ones — instead of the exact (no-FMA, Moeller-Trumbore) instantiation.
     * t/u/v then differ from the reference in the last bits; primitive ids and occlusion flags must not (tests/test_gpu_fast_hooks.py). */
    XRTG_FLAG_FAST_HOOK = 1u << 5,
    /* xrtg_trace_rays(any_hit = 1) with XRTG_FLAG_FAST_HOOK: out_hits[i].prim holds, ON ENTRY, the id of the primitive the shadow
     * ray starts on (-1 = unknown) — what the renderer knows when it traces an NEE ray and uses to pick the occluder section
     * of a small scene (hull pruning, small_scene.h). */
    XRTG_FLAG_HOOK_SRC_PRIM = 1u << 6
};

typedef struct xrtg_render_params {
    int32_t width, height;
    int32_t spp;           /* samples rendered by THIS call                                  */
    int32_t sample_offset; /* index of the first sample (spp split across GPUs / resumable)  */
    int32_t spp_total;     /* divisor of the final mean; 0 = spp (renderer.cpp:98)           */
    int32_t integrator;    /* xrtg_integrator                                                */
    int32_t max_depth;
    uint32_t seed;         /* counter RNG seed (ignored with XRTG_FLAG_EXACT)                */
    uint32_t flags;
    int32_t samples_per_wave; /* 0 = auto                                                    */
} xrtg_render_params;

typedef struct xrtg_stats {
    uint64_t samples;
    uint64_t closest_rays;   /* Scene::intersect calls the reference would make   */
    uint64_t shadow_rays;    /* Scene::occluded calls the reference would make    */
    uint64_t dropped_samples;/* NaN/inf/negative samples (renderer.cpp:57-73)     */
    uint64_t nodes_visited;  /* closest-hit BVH nodes fetched, with XRTG_FLAG_COUNTERS */
    uint64_t tris_tested;    /* closest-hit triangles tested, with XRTG_FLAG_COUNTERS */
    uint64_t nodes_visited_shadow; /* any-hit BVH nodes fetched, with XRTG_FLAG_COUNTERS */
    uint64_t tris_tested_shadow;   /* any-hit triangles tested, with XRTG_FLAG_COUNTERS */
    uint64_t tracking_steps; /* delta/ratio tracking steps (always counted)        */
    uint64_t kernel_launches;
    uint64_t extend_launches, shade_launches, connect_launches;
    float render_ms;         /* CUDA-event time of the device work                 */
    float extend_ms;         /* closest-hit traversal kernels (XRTG_FLAG_STAGE_TIMES or _COUNTERS) */
    float connect_ms;        /* any-hit traversal kernels                          */
    float shade_ms;
    float other_ms;
    float h2d_ms, d2h_ms;
    uint64_t primary_hits;   /* primary rays that hit something = entries of the compact bounce-0 queue */
    uint64_t bounce_entries; /* queue entries consumed by the fused per-bounce kernel of small scenes (0 = three-kernel pipeline) */
    uint64_t bounce_launches;/* launches of that kernel (counted in shade_launches as well) */
    uint64_t rays_traced;    /* rays the kernels actually traced: closest_rays + shadow_rays minus the primary samples the
                                screen-space scissor resolved without a ray (they still count as reference-equivalent rays) */
    uint64_t truncated_paths;/* volume paths cut by the iteration bound 4*max_depth+8 (capped at 4096) on medium crossings without a
                                scattering event; the reference's loop (integrator.h:418) is unbounded. 0 on every shipped scene */
    float reduce_ms;         /* multi-GPU scenes: the fused peer-memory reduce + finalize (CUDA events on device 0)            */
    int32_t n_devices;       /* devices that took part in the render                                                           */
    uint64_t untraced_closest; /* of closest_rays: counted as the reference's Scene::intersect calls but not traced — primary samples
                                  outside the screen-space scissor, extension rays of paths that lost the next Russian roulette     */
    uint64_t untraced_shadow;  /* of shadow_rays: NEE samples whose contribution is exactly zero (throughput instantiation)           */
} xrtg_stats;

/* Closest-hit record of the parity hooks. prim = global primitive id (-1 = miss). */
typedef struct xrtg_hit {
    float t, u, v;
    int32_t prim;
} xrtg_hit;

typedef struct xrtg_scene xrtg_scene;

typedef struct xrtg_scene_info {
    int32_t n_prims, n_triangles, n_bvh_nodes, bvh_depth;
    float bvh_sah_cost;
    float build_ms, upload_ms; /* whole host-side ingest (flatten + BVH) / H2D upload                    */
    uint64_t device_bytes; /* scene data resident in HBM      */
    uint64_t upload_bytes; /* bytes copied H2D by an upload    */
    float bvh_build_ms;    /* the BVH builder alone (host SAH, or the GPU LBVH kernels incl. their sort) */
    int32_t bvh_builder;   /* 0 = host binned SAH, 1 = GPU linear BVH, 2 = GPU PLOC (device ingest + build + collapse) */
    int32_t small_records_all, small_records_occ; /* plane-paired triangle records (80 B each) of a small scene's closest-hit /
                                                     occluder sections; 0 = no block (per-triangle lists or BVH only)  */
    int32_t small_flagged;  /* small scenes: primitives whose shadow rays can start behind a hull-pruned plane (they test the
                               unpruned occluder section)                                                               */
    int32_t n_wide_nodes;   /* deep scenes: nodes of the wide (4- or 8-child) tree the traversal kernel walks          */
    int32_t wide_arity;     /* 0 = none, 4 or 8                                                                         */
    int32_t n_devices;      /* devices holding a replica of the scene                                                    */
} xrtg_scene_info;

/* Development / test switches of the pipeline selection; -1 (or any negative value) = the measured default. None of them
 * changes WHAT is computed, only which kernels compute it — the tests flip them to compare one pipeline against another.
 * The render path reads no environment variable; XRT_TUNING="key=value,..." (keys = the field names) is parsed once, at scene
 * creation. */
typedef struct xrtg_tuning {
    int32_t fused_bounce;    /* small scenes: one k_bounce_small per bounce (1) or shade -> connect -> extend (0)              */
    int32_t volume_paths;    /* shallow BVHs: k_volume_paths (1) or the wavefront iterations (0)                              */
    int32_t scissor;         /* primary-ray screen-space scissor                                                              */
    int32_t brute_secondary, brute_shadow; /* small scenes: shared-memory triangle loops instead of the BVH walk              */
    int32_t thr_ext0, thr_ext, thr_con;    /* k_trace refill thresholds (0 = the simple run-to-completion kernels)            */
    int32_t steps_per_vote, leaf_threshold;
    int32_t thr_vol, spv_vol;              /* k_volume_paths lockstep-walk threshold / steps per vote                         */
    int32_t wide_bvh;        /* deep scenes: 8 = eight-child quantised nodes, 4 = four-child nodes, 2 = two-child nodes       */
    int32_t max_leaf;        /* creation time only (XRT_TUNING): triangles per leaf of the host SAH builder, 1..4             */
    int32_t workspace_mb;    /* byte budget of the per-wave queues (default 16384); small values force pixel-tiled waves       */
    int32_t stage_dump;      /* print every stage's CUDA-event time to stderr (with XRTG_FLAG_STAGE_TIMES)                    */
    int32_t primary_masks;   /* small scenes: screen-space candidate masks for the primary rays (1) or the BVH walk (0)       */
    int32_t gpu_build;       /* creation time only (XRT_TUNING): 1 = as if XRTG_BUILD_GPU were passed, 0 = never build on the device */
    int32_t ploc_radius;     /* creation time only: PLOC neighbour-search radius (default 16)                                 */
    int32_t ploc_ct_x16;     /* creation time only: SAH traversal-step cost of the PLOC leaf decision, in 1/16 (default 16)   */
    int32_t ploc_top;        /* creation time only: clusters at which PLOC hands over to the top-level sweep SAH (default 8) */
    int32_t grid_texture;    /* creation time only: density grids also live in a 3-D texture (1, default) that the throughput
                                instantiation samples with hardware trilinear filtering; 0 = global-memory lookups only              */
    int32_t overlap_connect; /* three-kernel pipeline: any hit of bounce b on a side stream, concurrently with the closest hit of bounce b+1 */
    int32_t ploc_weight;     /* creation time only: top-level sweep SAH weighs a side by its triangles (0) or clusters (1, default) */
} xrtg_tuning;

int xrtg_abi_version(void);
int xrtg_device_count(void);
const char* xrtg_last_error(void);

/* Copies the PODs, builds the BVH (host binned SAH; on the device from 65536 mesh triangles on, see XRTG_BUILD_GPU), uploads
 * everything to `device`. */
int xrtg_scene_create(const xrtg_scene_desc* desc, int device, xrtg_scene** out);

enum {
    /* Build the BVH ON THE GPU (linear BVH: Morton codes, radix sort, Karras radix tree, bottom-up fit) instead of the
     * host SAH builder: ~100x faster to build, a somewhat slower tree to traverse. Results are identical (any valid BVH
     * returns what the reference's brute-force loops return). */
    XRTG_BUILD_LBVH_GPU = 1u << 0,
    /* Ingest AND build on the device (csrc/gpu_build.cu): the raw triangle array is copied up once; the per-triangle records,
     * a PLOC tree (parallel locally-ordered clustering over 63-bit Morton order down to the last 8 clusters, SAH-decided leaves of
     * up to four triangles) and
     * its collapse into eight-child quantised nodes are produced in HBM. Tens of milliseconds for a million triangles instead of
     * the host path's ~0.5-1.3 s, and a tree that traverses like the host SAH one. No pinned staging copies are made until
     * xrtg_scene_upload or a multi-GPU replica asks for them. This is the DEFAULT for scenes of 65536 mesh triangles or more; the
     * flag lowers that threshold to 1024 (below it the host path is taken regardless: it is faster there). Inputs the clustering
     * cannot handle (a tree deeper than the traversal stacks) fall back to the host builder. */
    XRTG_BUILD_GPU = 1u << 1,
    /* Force the host binned-SAH builder, whatever the size of the scene. */
    XRTG_BUILD_HOST = 1u << 2
};
/* xrtg_scene_create with build flags. */
int xrtg_scene_create2(const xrtg_scene_desc* desc, int device, uint32_t build_flags, xrtg_scene** out);
/* Multi-GPU scene: the SAME host-side ingest and BVH build, then one replica of the scene arrays per device (devices[0..ngpus),
 * NULL = 0..ngpus-1). xrtg_render / xrtg_render_device on such a handle split the samples of the call across the devices
 * (device g renders sample indices [g*spp/G, (g+1)*spp/G) of every pixel; the counter RNG is keyed by sample index, so the
 * union is the 1-GPU sample set), one host thread and one stream per device, and finish with ONE fused kernel per device that
 * pulls its slice of every device's per-pixel SUM over NVLink peer memory, adds them in device order, applies the
 * reference's `image /= n_samples` (renderer.cpp:98) and stores the slice into device 0's image — reduce and finalize in
 * one pass, no NCCL, no Python. Single process; XRTG_FLAG_EXACT renders are not split (one mt19937 stream per pixel). */
int xrtg_scene_create_multi(const xrtg_scene_desc* desc, int ngpus, const int* devices, uint32_t build_flags, xrtg_scene** out);
/* Number of devices behind a handle (1 for xrtg_scene_create). */
int xrtg_scene_device_count(const xrtg_scene* scene);
int xrtg_scene_set_tuning(xrtg_scene* scene, const xrtg_tuning* tuning);
/* Every workspace buffer of the handle (ray / hit / shadow queues, radiance, counters, ...) is allocated with 256-byte guard bands
 * of a known pattern on both sides; this counts the guard bytes that were overwritten since allocation (0 = no kernel wrote out
 * of bounds). Synchronises the scene's stream(s). The GPU tests call it after every render (tests/conftest.py). */
int xrtg_scene_check_guards(xrtg_scene* scene, int* violations);

/* ---- one process per GPU (torchrun / MPI style deployments): the same fused reduce + finalize over CUDA IPC ------------- */
/* An exportable device buffer owned by the scene (plain cudaMalloc on the scene's device, so that cudaIpcGetMemHandle applies to
 * it; the pointer stays valid until a larger size is requested for the same slot or the scene is destroyed). Slot 0..3; by
 * convention slot 0 holds the rank's per-pixel SUM and slot 1, on the root rank, the final image. */
int xrtg_exchange_buffer(xrtg_scene* scene, int slot, size_t bytes, void** device_ptr);
#define XRTG_IPC_HANDLE_BYTES 64
int xrtg_ipc_export(const void* device_ptr, unsigned char handle[XRTG_IPC_HANDLE_BYTES]);
/* Maps a buffer exported by another process on another (peer-accessible) device into this process. */
int xrtg_ipc_open(xrtg_scene* scene, const unsigned char handle[XRTG_IPC_HANDLE_BYTES], void** device_ptr);
int xrtg_ipc_close(xrtg_scene* scene, void* device_ptr);
/* out[i] = (parts[0][i] + parts[1][i] + ... in this order) / divisor for i in [first, first + count); divisor <= 0 leaves
 * the sum. `parts` and `out` may be local, peer-device or IPC-mapped pointers; asynchronous on `cuda_stream`. Every rank
 * calls it on its own slice; the caller orders it after the renders of all ranks (a stream-ordered barrier). */
int xrtg_reduce_finalize(xrtg_scene* scene, const float* const* parts, int nparts, float* out, size_t first, size_t count,
                         float divisor, void* cuda_stream);

/* Re-copies the already-built scene arrays host(pinned)->device (the e2e H2D leg). */
int xrtg_scene_upload(xrtg_scene* scene);
int xrtg_scene_get_info(const xrtg_scene* scene, xrtg_scene_info* out);
void xrtg_scene_destroy(xrtg_scene* scene);

/* Render into a HOST buffer rgb[W*H*3]: mean radiance (or the sum with XRTG_FLAG_SUM_ONLY).
 * The timed D2H copy is part of the call. stats may be NULL. */
int xrtg_render(xrtg_scene* scene, const xrtg_camera* cam, const xrtg_render_params* p, float* rgb_host,
                xrtg_stats* stats);
/* Same, into a DEVICE buffer on the scene's device, asynchronously on `cuda_stream`
 * (a cudaStream_t passed as void*; NULL = the legacy default stream). The caller synchronises. */
int xrtg_render_device(xrtg_scene* scene, const xrtg_camera* cam, const xrtg_render_params* p,
                       float* rgb_device, void* cuda_stream, xrtg_stats* stats);

/* Parity hook: primary rays only. jitter_uv = W*H*spp*2 floats in [0,1) laid out [(i*W+j)*spp+k][2]
 * (host pointer) or NULL to take them from the pixel's mt19937 stream as renderer.cpp:44-47 does.
 * out = W*H*spp hits (host). flags: XRTG_FLAG_BRUTE_FORCE honoured. */
int xrtg_trace_primary(xrtg_scene* scene, const xrtg_camera* cam, int width, int height, int spp,
                       const float* jitter_uv, uint32_t flags, xrtg_hit* out);

/* Parity hook: arbitrary rays. org/dir = n*3 floats (host), tmax = n floats or NULL (= +inf).
 * any_hit=0: Scene::intersect semantics -> out_hits[n].  any_hit=1: Scene::occluded(ray,tmax)
 * semantics (emitter proxies skipped) -> out_hits[i].prim = 0/1 occluded flag in prim>=0. */
int xrtg_trace_rays(xrtg_scene* scene, int64_t n, const float* org, const float* dir, const float* tmax,
                    int any_hit, uint32_t flags, xrtg_hit* out_hits);

/* Structural check of the acceleration structures RESIDENT ON THE DEVICE, whichever builder produced them: the two-child tree, the
 * leaf-ordered triangles, the eight-child quantised tree and its node-ordered records are copied back and walked on the host —
 * every triangle in exactly one leaf, every (decoded, margin-shrunk) child box contains the triangles below it, depths fit the
 * traversal stacks. *n_errors = number of violations (xrtg_last_error() describes the first). Scenes of fewer than two mesh
 * triangles have nothing to check. */
int xrtg_scene_selfcheck(xrtg_scene* scene, int* n_errors);

/* Host-only structural check of the SAH BVH builder (no CUDA device needed): builds the tree over n triangles (9 floats each:
 * v0 v1 v2) and verifies that every triangle is referenced by exactly one leaf, that every child box (minus the conservative
 * padding) contains its subtree, that leaves hold at most max_leaf triangles and that the depth fits the traversal stacks;
 * then collapses the tree to the four-child form deep scenes are traversed in and checks the same properties on it.
 * Returns 0 if the tree is valid, a negative xrtg_status otherwise; the out parameters may be NULL. */
int xrtg_bvh_selftest(const float* tris9, int n, int max_leaf, int* n_nodes, int* depth, float* sah_cost);

/* Host-only structural check of the sweep-SAH builder that splits the top levels of a device-built tree (csrc/gpu_build.cu:
 * TopBuilder; no CUDA device needed): n clusters given by their boxes (lo3 / hi3 = n * 3 floats) and triangle counts (NULL = 1
 * each) — every cluster referenced exactly once, every node's box the union of its children's, counts adding up. *depth = depth of
 * the tree (coincident boxes must give a logarithmic depth: equal-cost splits take the most balanced one). */
int xrtg_top_sah_selftest(const float* lo3, const float* hi3, const uint32_t* counts, int n, int by_clusters, int* depth);

/* Host-only check of the plane-paired triangle block that scenes of at most 64 triangles are traced from (no CUDA device
 * needed): builds the block over n triangles (9 floats each; emitter_flags[i] & 1 marks emitter proxies, may be NULL) and
 * verifies that every triangle sits in exactly one record of the closest-hit section, every non-emitter in exactly one
 * record of the occluder section, and that each record's plane and barycentric equations reproduce the triangle it came from.
 * Returns 0 for a valid block, 1 if the scene has too few coplanar pairs for a block to pay (none is built, the per-triangle
 * lists are used), a negative xrtg_status on an inconsistent block. Out parameters may be NULL. */
int xrtg_small_scene_selftest(const float* tris9, const int* emitter_flags, int n, int* n_records_all, int* n_records_occ, int* n_planes);

/* Image post of the reference's Image class on the device (image.h:80-136): optional gammaCorrection — pow(x, 1/gamma) per
 * channel (image.h:80-90), gamma <= 0 skips it — followed by the 8-bit quantisation shared by writePPM and writeMat,
 * clamp(uint32(255*x), 0, 255), written RGB (PPM order, bgr = 0) or BGR (cv::Mat order, bgr = 1). rgb_host = W*H*3 floats,
 * out_host = W*H*3 bytes (both host pointers). */
int xrtg_image_to_u8(int device, const float* rgb_host, int width, int height, float gamma, int bgr, uint8_t* out_host);
/* xrtg_render followed by that image post WITHOUT the float image leaving HBM: render, `image /= n_samples`, gammaCorrection
 * and the 8-bit quantisation run on the device and only width*height*3 BYTES are copied to the host (what the reference's
 * example mains do after render(): image.gammaCorrection(1.2f); image.writePPM(...), cornellbox.cpp:65-75). */
int xrtg_render_u8(xrtg_scene* scene, const xrtg_camera* cam, const xrtg_render_params* p, float gamma, int bgr, uint8_t* out_host,
                   xrtg_stats* stats);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* XRTGPU_H */
