/*
 * b200rt.h — C ABI of the B200-native rendering core that sits behind the
 * ceng795 ray tracer's `Scene::render_image` seam.
 *
 * The reference has no FFI; its only stable seam is the set of `Scene` member
 * functions that `main.cpp` fans out over std::threads:
 *   HW1      void Scene::render_image(int, Vector3i*, int, int) const   HW1/Scene.h:34-35
 *   HW2..HW7 void Scene::render_image(int, Pixel*,    int, int) const   HW7/include/Scene.h:59-60
 *   PPM      reset_hash_grid / eye_trace_lines / build_hash_grid /
 *            trace_n_photons / density_estimation                      PPM/include/Scene.h:48-55
 *
 * Every entry point below replaces one of those calls (cited per function).
 * Plain pointers and sizes only; no C++ or torch types cross this boundary.
 * All functions return 0 on success or a negative B200RT_E_* code; they never
 * throw.  `b200rt_last_error()` returns a thread-local message.
 *
 * Two layers live in this header:
 *   1. b200rt_builder_*  — host-side scene assembly.  Restates (in C++, IEEE
 *      binary32, no FMA contraction) the host arithmetic the reference does at
 *      load time: transformation matrices, camera frames, triangle records,
 *      vertex normals and — crucially — the reference's BVH *topology*
 *      (HW7/src/Bounding_volume_hierarchy.cpp:3-29) so that the flattened
 *      tree's DFS leaf order equals the reference's tie-breaking order.
 *      INTEGRATION.md shows the alternative producer of the same B200SceneDesc:
 *      a flatten.cpp TU compiled inside the reference tree.
 *   2. b200rt_scene_* / b200rt_render* / b200rt_ppm_* — the device side.
 */
#ifndef B200RT_H_
#define B200RT_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes ------------------------------------------------------ */
#define B200RT_OK 0
#define B200RT_E_INVALID (-1)   /* bad argument / malformed description      */
#define B200RT_E_CUDA (-2)      /* CUDA runtime error (see last_error)        */
#define B200RT_E_NOGPU (-3)     /* no CUDA device: there is NO CPU fallback   */
#define B200RT_E_UNSUPPORTED (-4)
#define B200RT_E_NOMEM (-6)     /* host allocation failed inside the library (never thrown across the ABI) */
#define B200RT_E_SINGULAR (-5)  /* reference exit(-1): "NOT INVERTIBLE MATRIX" HW3/src/Transformation.cpp:78-82 */

/* ---- per-tree semantics -------------------------- */
enum b200rt_semantics {
  B200RT_SEM_HW1 = 1,
  B200RT_SEM_HW2 = 2,
  B200RT_SEM_HW3 = 3,
  B200RT_SEM_HW4 = 4,
  B200RT_SEM_HW5 = 5,
  B200RT_SEM_HW6 = 6,
  B200RT_SEM_HW7 = 7,
  B200RT_SEM_PPM = 8
};

/* ---- BVH child references ----------------------------------------------- */
#define B200RT_KIND_INNER 0u    /* index into nodes[]                          */
#define B200RT_KIND_TRI 1u      /* mesh triangle: index into tri_* arrays      */
#define B200RT_KIND_SPHERE 2u   /* index into spheres[]                        */
#define B200RT_KIND_TRIPRIM 3u  /* stand-alone Triangle with own transform     */
#define B200RT_KIND_INSTANCE 4u /* index into instances[]                      */
#define B200RT_KIND_LEAF1 5u    /* fast-BVH leaf of 1..4 triangles (kinds 5..8):  */
#define B200RT_KIND_LEAF4 8u    /*   index = first entry in fast_tri_* arrays     */
#define B200RT_KIND_FNODE 9u    /* index into fast_nodes[]                        */
#define B200RT_KIND_LIGHTINST 10u /* HW7 Light_mesh: index into instances[], entered WITHOUT a gate box */
#define B200RT_KIND_NONE 15u
#define B200RT_REF(kind, idx) ((((uint32_t)(kind)) << 28) | ((uint32_t)(idx)))
#define B200RT_REF_KIND(r) ((r) >> 28)
#define B200RT_REF_INDEX(r) ((r)&0x0FFFFFFFu)
#define B200RT_REF_NONE B200RT_REF(B200RT_KIND_NONE, 0)

/* One inner node = the reference's `BVH` object (HW7 72 B pointer node) with
 * its two children's gate boxes pulled in, 64 B, 64-byte aligned.  The box of
 * a child is only meaningful (and only tested) for INNER and INSTANCE
 * children: the reference gates exactly those with Bounding_box::intersect
 * (BVH::intersect HW3/src/Bounding_volume_hierarchy.cpp:32-35,
 *  Mesh_instance::intersect HW3/include/Mesh.h:59-62); triangles and spheres
 * are tested unconditionally once their parent is visited. */
typedef struct b200rt_bvh_node {
  float lmin[3], lmax[3];
  float rmin[3], rmax[3];
  uint32_t left, right;
  uint32_t pad[2];
} b200rt_bvh_node;

typedef struct b200rt_mesh {
  uint32_t root;      /* child ref of Mesh::bvh (INNER, or TRI when 1 triangle) */
  float root_box[6];  /* min xyz, max xyz of the root (gate when INNER)         */
  int32_t tri_begin, tri_count; /* triangles [begin,begin+count) in DFS order  */
  int32_t node_begin, node_count;
  uint32_t fast_root; /* FNODE / LEAFn ref into the SAH tree the device walks, or NONE */
  int32_t pad;
} b200rt_mesh;

#define B200RT_INST_MOVING 1u     /* velocity != 0: per-ray inverse (HW3/include/Mesh.h:76-95) */
#define B200RT_INST_DIRECT 2u     /* HW2: Mesh sits in the top BVH itself, no gate / transform  */
#define B200RT_INST_REFRACTIVE 4u /* HW7: culling off inside glass (HW7/include/Mesh.h:73-75)   */
#define B200RT_INST_LIGHT 8u      /* HW7 Light_mesh: no own gate box, emissive (HW7/include/Light_mesh.h:52-68) */
typedef struct b200rt_instance {
  float inv[12];      /* rows 0..2 of M^-1 (point: + column 3)                */
  float nrm[9];       /* upper 3x3 of transpose(M^-1)                         */
  float fwd[16];      /* M, only used when MOVING                              */
  float box[6];       /* world-space gate box                                 */
  float velocity[3];
  int32_t mesh;
  int32_t material;
  int32_t texture;
  uint32_t flags;
  int32_t rank;       /* DFS position among top-level objects (tie-break)     */
  int32_t light;      /* LIGHT: index into lights[] (radiance), else -1        */
  int32_t pad[2];
} b200rt_instance;

#define B200RT_XF_IDENTITY 1u
#define B200RT_XF_MOVING 2u
#define B200RT_XF_LIGHT 4u        /* HW7 Light_sphere (HW7/include/Light_sphere.h:16-24) */
typedef struct b200rt_sphere {
  float center[3];
  float radius;
  float inv[12];
  float nrm[9];
  float fwd[16];
  float velocity[3];
  int32_t material;
  int32_t texture;
  uint32_t flags;
  int32_t rank;
  int32_t light;      /* LIGHT: index into lights[], else -1 */
  int32_t pad;
} b200rt_sphere;

typedef struct b200rt_triprim {
  float v0[3], a1[3], a2[3], normal[3]; /* a1 = v0-v1, a2 = v0-v2 (object space) */
  float inv[12];
  float nrm[9];
  int32_t material;
  uint32_t flags;
  int32_t rank;
} b200rt_triprim;

/* HW4+ textures (HW7/include/Texture.h:40-55).  Image texels are floats 0..255
 * (after the optional degamma, HW7/src/Texture.cpp:13-24), 3 per texel, in
 * `texels` starting at texel_offset (in floats). */
enum b200rt_tex_kind { B200RT_TEX_IMAGE = 0, B200RT_TEX_PERLIN = 1 };
enum b200rt_tex_interp { B200RT_TEX_NEAREST = 0, B200RT_TEX_BILINEAR = 1 };
enum b200rt_tex_decal { B200RT_TEX_REPLACE_KD = 0, B200RT_TEX_BLEND_KD = 1, B200RT_TEX_REPLACE_ALL = 2 };
enum b200rt_tex_appearance { B200RT_TEX_REPEAT = 0, B200RT_TEX_CLAMP = 1, B200RT_TEX_VEIN = 2, B200RT_TEX_PATCH = 3 };
typedef struct b200rt_texture {
  int32_t kind, interp, decal, appearance;
  float normalizer, scaling, bump_multiplier;
  int32_t is_bump;
  int32_t width, height;
  int64_t texel_offset;
  int32_t perm[16];   /* Perlin permutation of the HOST object (the reference shuffles with a
                         clock-seeded mt19937, HW7/src/Perlin_noise.cpp:17-20)              */
} b200rt_texture;

/* HW6+ BRDFs (HW7/src/*_BRDF.cpp) */
enum b200rt_brdf_kind {
  B200RT_BRDF_PHONG = 0, B200RT_BRDF_MODIFIED_PHONG = 1, B200RT_BRDF_BLINN_PHONG = 2,
  B200RT_BRDF_MODIFIED_BLINN_PHONG = 3, B200RT_BRDF_TORRANCE_SPARROW = 4
};
typedef struct b200rt_brdf {
  int32_t kind;
  float exponent;
  int32_t normalized;
  float refractive_index;
} b200rt_brdf;

typedef struct b200rt_material {
  float ambient[3], diffuse[3], specular[3], mirror[3], transparency[3];
  float roughness, refraction_index, phong_exponent;
  int32_t brdf;      /* -1 = built-in Blinn-Phong                               */
  int32_t degamma;
  float pad[4];
} b200rt_material;  /* 96 B */

enum b200rt_light_kind {
  B200RT_LIGHT_POINT = 0,
  B200RT_LIGHT_AREA = 1,
  B200RT_LIGHT_SPOT = 2,
  B200RT_LIGHT_DIRECTIONAL = 3,
  B200RT_LIGHT_ENV = 4,
  B200RT_LIGHT_MESH = 5,
  B200RT_LIGHT_SPHERE = 6
};
typedef struct b200rt_light {
  int32_t kind;
  float position[3];
  float intensity[3];
  float edge1[3], edge2[3], normal[3]; /* area; spot: edge1 = direction, edge2 = {cos(cov/2), cos(fall/2), 0} */
  int32_t object;                      /* mesh light: instance index; sphere light: sphere index; env: texture index */
  int32_t cdf_begin, cdf_count;        /* mesh light: range in light_cdf[] / light_tri[] (HW7/include/Light_mesh.h:19-50) */
  float total_area;
  float pad;
} b200rt_light;

typedef struct b200rt_camera {
  float e[3], u[3], v[3], w[3];
  float top_left[3], su[3], sv[3];       /* Camera ctor HW3/include/Camera.h:17-34 */
  float left, right, bottom, top, distance; /* image plane as used (after DOF rescale) */
  int32_t width, height;
  int32_t samples_sqrt;                  /* (int)sqrt(NumSamples), min 1            */
  float aperture;
  int32_t pad[3];
} b200rt_camera;

typedef struct B200SceneDesc {
  int32_t semantics;
  int32_t max_depth;
  float background[3];
  float ambient[3];
  float shadow_eps;
  float isect_eps;     /* kEpsilon / intersection_test_epsilon: 1e-6 (PPM 1e-4) */

  uint32_t root;       /* Scene::bvh: INNER, a single object, or NONE          */
  float root_box[6];

  int32_t n_nodes;
  const b200rt_bvh_node* nodes;

  int32_t n_tris;
  const float* tri_geom;    /* 12 floats per triangle: v0, a1=v0-v1, a2=v0-v2, face normal */
  const int32_t* tri_vidx;  /* 3 per triangle: indices into vert_* (smooth shading, uv)     */
  const int32_t* tri_info;  /* 2 per triangle: material, flags (bit0 = smooth)             */

  /* Device-oriented SAH BVH over the same triangles (csrc/fast_bvh.h).  A triangle
   * found through it is admitted only if the reference's innermost gate for it
   * (tri_gate: the box of the reference BVH node that holds it) passes
   * Bounding_box::intersect — all outer gates are supersets and pass with it. */
  int32_t n_fast_nodes;
  const b200rt_bvh_node* fast_nodes;
  int32_t n_fast_tris;         /* triangles of meshes with >= 2 triangles        */
  const float* fast_tri_geom;  /* 12 floats per triangle, leaf order             */
  const int32_t* fast_tri_ref; /* leaf-order position -> triangle index above    */
  const float* tri_gate;       /* 6 floats per triangle (DFS order): min, max    */

  int32_t n_verts;
  const float* vert_pos;    /* 3 per vertex */
  const float* vert_normal; /* 3 per vertex, finalised (HW3/src/Scene.cpp:872-877) */

  int32_t n_meshes;
  const b200rt_mesh* meshes;
  int32_t n_instances;
  const b200rt_instance* instances;
  int32_t n_spheres;
  const b200rt_sphere* spheres;
  int32_t n_triprims;
  const b200rt_triprim* triprims;
  int32_t n_materials;
  const b200rt_material* materials;
  int32_t n_lights;
  const b200rt_light* lights;
  int32_t n_cameras;
  const b200rt_camera* cameras;

  /* ---- HW4..HW7 ---- */
  const float* tri_uv;         /* 6 floats per triangle: uv of the three vertices (0 if the mesh has no image texture) */
  const int32_t* tri_tex;      /* 1 per triangle: texture id of its mesh, -1 none (HW7/src/Mesh_triangle.cpp:107) */
  int32_t n_textures;
  const b200rt_texture* textures;
  int64_t n_texels;            /* floats */
  const float* texels;
  int32_t n_brdfs;
  const b200rt_brdf* brdfs;
  int32_t n_light_cdf;
  const float* light_cdf;      /* cumulative area / total area */
  const int32_t* light_tri;    /* triangle index per cdf entry (file order of the LightMesh faces) */
  int32_t background_texture;  /* texture id or -1 (HW7/src/Scene.cpp:476-484) */
  int32_t env_light;           /* index into lights[] of the SphericalDirectionalLight or -1 */
  int32_t integrator;          /* 0 ray tracing, 1 path tracing (HW7/src/Scene.cpp:1573-1582) */
  int32_t importance_sampling; /* path tracing: 1 = cosine sampling */
} B200SceneDesc;

/* ======================================================================== */
/* 1. Host-side builder (restates the reference's load-time arithmetic)      */
/* ======================================================================== */
typedef struct b200rt_builder b200rt_builder;

/* Transformation helpers: row-major 4x4, same arithmetic as
 * Scaling/Translation/Rotation ctors (HW3/src/Transformation.cpp:6-70) and
 * Matrix4x4::operator* (HW3/src/Matrix4x4.cpp:52-62). angle in DEGREES, the
 * helper applies the reference's `angle*degree_to_pi` (HW3/src/Scene.cpp:322,615). */
void b200rt_xf_identity(float out[16]);
void b200rt_xf_scaling(float x, float y, float z, float out[16]);
void b200rt_xf_translation(float x, float y, float z, float out[16]);
void b200rt_xf_rotation(float angle_deg, float x, float y, float z, float out[16]);
void b200rt_xf_mul(const float a[16], const float b[16], float out[16]);
int b200rt_xf_invert(const float m[16], float out[16]); /* Matrix4x4::invert_matrix */

typedef struct b200rt_camera_in {
  float position[3], gaze[3], up[3];
  float near_plane[4]; /* l r b t */
  float near_distance;
  int32_t width, height;
  int32_t num_samples;     /* as written in the XML; builder applies (int)sqrt */
  float focus_distance, aperture_size;
  int32_t left_handed;     /* HW7 handedness="left" (HW7/include/Camera.h:24-26) */
  int32_t simple;          /* HW7 type="simple": `gaze` holds the gaze POINT, near plane from fov_y (HW7/src/Scene.cpp:590-607) */
  float fov_y;             /* degrees */
} b200rt_camera_in;

b200rt_builder* b200rt_builder_create(int semantics);
void b200rt_builder_destroy(b200rt_builder*);
int b200rt_builder_set_globals(b200rt_builder*, const float background[3], const float ambient[3],
                               float shadow_eps, int max_depth);
int b200rt_builder_add_camera(b200rt_builder*, const b200rt_camera_in*);
int b200rt_builder_add_material(b200rt_builder*, const b200rt_material*);
int b200rt_builder_add_light(b200rt_builder*, const b200rt_light*);
int b200rt_builder_set_vertices(b200rt_builder*, const float* xyz, int n);
/* faces: 0-based indices, vertex_offset already added. Returns mesh id (>=0).
 * HW3+: also appends the implicit base instance (HW3/src/Scene.cpp:703-707). */
int b200rt_builder_add_mesh(b200rt_builder*, int material, int smooth, const int32_t* faces,
                            int n_faces, const float base_xf[16], const float velocity[3]);
int b200rt_builder_add_instance(b200rt_builder*, int base_mesh, int material,
                                const float xf[16], const float velocity[3]);
int b200rt_builder_add_triangle(b200rt_builder*, int material, const int32_t idx[3],
                                const float xf[16]);
int b200rt_builder_add_sphere(b200rt_builder*, int material, int center_vertex, float radius,
                              const float xf[16], const float velocity[3]);
/* ---- HW4..HW7 additions (ignored by earlier semantics) ---- */
int b200rt_builder_set_texcoords(b200rt_builder*, const float* uv, int n);
/* texels: width*height*3 floats 0..255 already degamma'd (NULL for Perlin). Returns texture id. */
int b200rt_builder_add_texture(b200rt_builder*, const b200rt_texture*, const float* texels);
int b200rt_builder_add_brdf(b200rt_builder*, const b200rt_brdf*);
/* Texture / texcoord offset of the mesh added last (HW7/src/Scene.cpp:1088-1098). */
int b200rt_builder_set_mesh_texture(b200rt_builder*, int mesh, int texture, int texture_offset);
int b200rt_builder_set_sphere_texture(b200rt_builder*, int sphere, int texture);
/* LightMesh (HW7/src/Scene.cpp:1246-1311): faces 0-based with the vertex offset applied. */
int b200rt_builder_add_light_mesh(b200rt_builder*, int material, const int32_t* faces, int n_faces,
                                  const float xf[16], const float radiance[3]);
int b200rt_builder_add_light_sphere(b200rt_builder*, int material, int center_vertex, float radius,
                                    const float xf[16], const float radiance[3]);
int b200rt_builder_set_integrator(b200rt_builder*, int path_tracing, int importance_sampling);
int b200rt_builder_set_background_texture(b200rt_builder*, int texture);

/* Builds per-mesh and top-level BVHs in the reference's topology, finalises
 * vertex normals and flattens.  The returned desc is owned by the builder. */
int b200rt_builder_finish(b200rt_builder*, const B200SceneDesc** out);

/* ======================================================================== */
/* 2. Device side                                                            */
/* ======================================================================== */
typedef struct b200rt_scene b200rt_scene;

typedef struct B200RenderOpts {
  int32_t row_begin, row_end; /* render rows [begin,end) (0,0 = all) : multi-GPU tile split */
  int32_t row_stride;         /* 1 = contiguous; >1 = reference-style interleave `j += height_increase`: groups of
                                 row_group rows start at row_begin, row_begin + row_stride, ... */
  uint64_t seed;              /* counter-based RNG key for stochastic configs */
  int32_t sample_begin, sample_end; /* strata subset [begin,end) of samples_sqrt^2 (0,0 = all) */
  int32_t exact_traversal;    /* 1 = visit every gated node like the reference (no t-pruning) */
  int32_t contexts;           /* row bands of the frame rendered concurrently on streams of their own (0 = the scene's
                                 default, 1 = one serial launch train); the result does not depend on it */
  int32_t row_group;          /* consecutive rows per interleave step (0 / 1 = single rows; b200rt_render_multi uses
                                 bands of 8 rows dealt round-robin to the devices) */
  int32_t pad[3];
} B200RenderOpts;

typedef struct B200RenderStats {
  uint64_t rays_primary, rays_shadow, rays_secondary;
  /* closest-hit (extend) launches only: */
  uint64_t nodes_visited, tris_tested, spheres_tested, instances_entered;
  /* any-hit (shadow) launches only: */
  uint64_t sh_nodes_visited, sh_tris_tested, sh_spheres_tested, sh_instances_entered;
  float ms_generate, ms_extend, ms_shade, ms_shadow, ms_total;
  int32_t waves, launches;
  int32_t node_bytes;   /* bytes of one nodes_visited unit inside a mesh: 80 (8-wide tree) or 64 (binary tree) */
  int32_t top_objects;  /* gate boxes of the flat top-level list (32 B each); every ray tests all of them and they are
                           counted in nodes_visited.  0: the top-level tree is walked instead */
  int32_t contexts;     /* row bands the frame was cut into; > 1: the ms_* stage times are sums over concurrent streams */
  int32_t pad;
} B200RenderStats;

int b200rt_device_count(void);
/* Copies the description into HBM on `device`; the caller keeps its host data. */
int b200rt_scene_create(const B200SceneDesc*, int device, b200rt_scene** out);
void b200rt_scene_destroy(b200rt_scene*);

/* Replaces the thread fan-out over Scene::render_image (HW7/src/main.cpp:35-46):
 * renders camera `camera_index` and writes W*H float4 {sum(color*w).rgb, sum(w)}
 * — the contents of the reference's Pixel{color, weight} — into caller-owned
 * HOST memory. */
int b200rt_render(b200rt_scene*, int camera_index, const B200RenderOpts*, float* rgbw_out,
                  B200RenderStats* stats_out);
/* Pipelined form for a host that renders frame after frame: returns once the frame is complete in HBM and its copy
 * into rgbw_out has been QUEUED (staged device-to-device, then copied on a stream of its own), so the copy crosses PCIe
 * while the next b200rt_render* call of the scene computes.  rgbw_out should be page-locked for the overlap to happen
 * and must stay valid until b200rt_render_wait, which blocks until every queued copy has landed.  At most two copies
 * are in flight; a third call waits for the oldest. */
int b200rt_render_async(b200rt_scene*, int camera_index, const B200RenderOpts*, float* rgbw_out,
                        B200RenderStats* stats_out);
int b200rt_render_wait(b200rt_scene*);
/* Same, but the result stays in HBM (device pointer returned, owned by scene). */
int b200rt_render_device(b200rt_scene*, int camera_index, const B200RenderOpts*,
                         float** d_rgbw_out, B200RenderStats* stats_out);
/* Parity hook: closest primary hit per pixel. ids_out[2*p] = top-level object
 * rank (DFS position in Scene::bvh, -1 = miss), ids_out[2*p+1] = triangle index
 * inside the mesh's DFS order (or -1); t_out[p] = hit t (inf on miss). */
int b200rt_primary_ids(b200rt_scene*, int camera_index, const B200RenderOpts*, int32_t* ids_out,
                       float* t_out);
/* Generic closest-hit query on caller-provided rays (host arrays, 8 floats per
 * ray: o.xyz, time, d.xyz, tmax-unused).  Test/measurement hook for a4-a8. */
int b200rt_trace_rays(b200rt_scene*, const float* rays, int64_t n, int32_t exact_traversal,
                      int32_t* ids_out, float* t_out, B200RenderStats* stats_out);

/* ---- PPM: progressive photon mapping (PPM/include/Scene.h:48-55, driven by PPM/src/main.cpp:33-102) ----
 * The reference updates a hit point under its mutex, one accepted photon at a time:
 *   g = (n*a + a)/(n*a + 1);  r2 *= g;  n++;  tau = (tau + f*phi) * g         (PPM/src/Scene.cpp:135-169)
 * so r2 is a function of n alone, r2(n) = r0^2 * R(n) with R(n) = prod_{k<n} g_k, and
 *   tau_final = R(n_final) * sum_k f_k*phi_k / R(k).
 * The device keeps n (claimed with atomicCAS after the radius test against r2(n)) and S = sum f*phi/R(k)
 * (atomicAdd): the same per-photon estimator, lock-free, in some serialisation order like the reference's
 * threads.  density_estimation's tau/(pi r2 N) is then S/(pi r0^2 N): additive across GPUs. */
typedef struct B200PpmOpts {
  int64_t photons;        /* photons to emit in this call                                   */
  uint64_t first_photon;  /* global index of the first one: RNG stream id (sharding / calls) */
  uint64_t seed;
  int32_t pad[4];
} B200PpmOpts;
typedef struct B200PpmStats {
  uint64_t hit_points, grid_entries, photons, photon_rays, cell_scans, deposits;
  uint64_t nodes_visited, tris_tested;   /* photon-pass traversal counters */
  float ms_eye, ms_grid, ms_photon, ms_extend, ms_deposit, ms_density;
  float initial_radius;
  int32_t launches;
  int32_t node_bytes, top_objects;
  /* the scan grid of the photon pass (rebuilt as the radii shrink; the ACCEPTED pairs are the reference's either way):
   * cell_scans above is what the reference's own loop over its r0 grid would have visited, entries_scanned what the
   * device tested, scan_cell the width of a cell of the grid in use (2 * initial_radius = the reference's), grid_rebuilds
   * the rebuilds this call did */
  uint64_t entries_scanned;
  float scan_cell;
  int32_t grid_rebuilds;
} B200PpmStats;
/* reset_hash_grid + eye_trace_lines + build_hash_grid for one camera */
int b200rt_ppm_eye_pass(b200rt_scene*, int camera_index, uint64_t seed, B200PpmStats*);
/* trace_n_photons: emits opts->photons photons from lights[0] and deposits them */
int b200rt_ppm_trace_photons(b200rt_scene*, const B200PpmOpts*, B200PpmStats*);
/* density_estimation(pixels, total_num_of_photons): W*H float4 {sum c*w, sum w} into host memory */
int b200rt_ppm_density(b200rt_scene*, double total_num_of_photons, float* rgbw_out);
/* multi-GPU: device float4 per hit point {S.rgb, n}; all-reduce(sum) it across ranks and feed it back */
int b200rt_ppm_export(b200rt_scene*, float** d_S_n, int64_t* n_hit_points);
int b200rt_ppm_density_from(b200rt_scene*, const float* d_S_n, double total_num_of_photons, float* rgbw_out);
/* test hook: 16 floats per hit point {position, normal, w_o, attenuation, material, pixel, pixel_weight, n} */
int b200rt_ppm_get_hit_points(b200rt_scene*, float* out16, int64_t capacity, int64_t* n_out);

/* ---- multi-GPU -------------------------------------------------------------------------------
 * The reference fans the rows of a frame out to its threads, `j = i; j += thread_count`, all of them writing one
 * shared Pixel array (HW7/src/main.cpp:35-46, HW7/src/Scene.cpp:28,41), and hands photon_count / thread_count
 * photons of a pass to each thread, all of them updating one shared Hit_point array (PPM/src/main.cpp:74-98,
 * PPM/src/Scene.cpp:135-169).  Here the scene is replicated, one b200rt_scene per GPU, and
 *   - an image is split into bands of 8 rows dealt round-robin to the GPUs (1-sample cameras: the final combine
 *     kernel of every GPU stores its pixels straight into the owner GPU's Pixel array over NVLink peer memory) or into
 *     batches of sample strata (multi-sample / splatting cameras: each GPU accumulates locally, then adds its frame
 *     into the owner's with 16-byte vector atomics over NVLink);
 *   - completion travels the same way: a system-scope counter at the end of the shared frame, bumped by each GPU's
 *     stream after its stores and awaited by a kernel on the owner's stream.  No host thread, no library collective;
 *   - a PPM pass is split by photon index (every photon is its own RNG stream) and the accepted-photon counts n are
 *     summed over the GPUs after each pass, so every GPU tests the next pass against the same shared radii r2(n).
 * Two ways in: b200rt_render_multi / b200rt_ppm_*_multi for ONE host thread (process) that owns all the GPUs — what the
 * reference's main would call — and the export / attach primitives underneath for one process per GPU (the handle is
 * plain bytes: send it through any channel; bench.py uses torch.distributed). */
typedef struct B200SharedFrame {
  uint8_t ipc_handle[64];   /* cudaIpcMemHandle_t of the owner's Pixel array (other processes)          */
  uint64_t device_ptr;      /* the same array as a device pointer (threads of the exporting process)    */
  int64_t pid;              /* exporting process                                                          */
  uint64_t tail_offset;     /* bytes from the array to its hand-shake counters (csrc/multi_gpu.cuh: FrameTail) */
  int32_t device, width, height;
  int32_t direct;           /* 1: contributors store pixels in place (1-sample cameras); 0: they add     */
} B200SharedFrame;
/* Owner: (re)allocates the Pixel array of the camera and describes it. */
int b200rt_frame_export(b200rt_scene* owner, int camera_index, B200SharedFrame* out);
/* Contributor (the owner itself included): later b200rt_render_shared calls of this scene go to that frame. */
int b200rt_frame_attach(b200rt_scene*, const B200SharedFrame*);
int b200rt_frame_detach(b200rt_scene*);
/* Owner, once per frame and before anybody's b200rt_render_shared of that frame: opens frame k = 1, 2, ... on the
 * owner's stream (clears the array unless `direct`), after the previous frame has been collected. */
int b200rt_frame_begin(b200rt_scene* owner);
/* Every attached scene, exactly once per frame: renders its share (opts: row_begin / row_stride / row_group or
 * sample_begin / sample_end) and delivers it into the shared frame, waiting on the device for frame_begin. */
int b200rt_render_shared(b200rt_scene*, int camera_index, const B200RenderOpts*, B200RenderStats* stats_out);
/* Owner: waits (on the device) until `parts` shares of the current frame have arrived, then copies the frame to host
 * memory (rgbw_out NULL: it stays in HBM; d_rgbw_out optional). */
int b200rt_frame_collect(b200rt_scene* owner, int parts, float* rgbw_out, float** d_rgbw_out);
/* As b200rt_render_async: waits for the shares on the device, queues the host copy, returns; b200rt_render_wait. */
int b200rt_frame_collect_async(b200rt_scene* owner, int parts, float* rgbw_out);
/* One host thread, n scenes of the same description on n devices (scenes[0] owns the frame): export + attach + begin +
 * one worker thread per device for render_shared + collect.  The frame equals b200rt_render's on one device: bit for
 * bit for 1-sample cameras, up to float summation order for splatting / multi-sample cameras (as the reference's own
 * threads).  stats_out: sums over the devices; ms_total = the frame's latency on the owner's stream.  If a share
 * fails (or never arrives: the wait on the device gives up after 20 s) the call returns its error and the next call
 * exports and attaches the frame afresh; a process-per-GPU host does the same after an error from
 * b200rt_render_shared / b200rt_frame_collect (b200rt_frame_export, then b200rt_frame_attach on every rank). */
int b200rt_render_multi(b200rt_scene* const* scenes, int n_scenes, int camera_index, const B200RenderOpts*,
                        float* rgbw_out, B200RenderStats* stats_out);
/* Share of device `part` of `parts` for camera `camera_index`, as b200rt_render_multi deals it (exposed so a
 * process-per-GPU host uses the same split). */
int b200rt_multi_share(b200rt_scene*, int camera_index, int part, int parts, B200RenderOpts* opts_inout);

/* PPM across GPUs.  Every GPU runs b200rt_ppm_eye_pass itself (hit points are numbered canonically, so index i is the
 * same hit point everywhere).  After each pass's b200rt_ppm_trace_photons:
 *   b200rt_ppm_counts_delta  device uint32 per hit point: photons accepted by THIS GPU since the last sync
 *   (sum it over the GPUs in place: an NCCL all-reduce between processes; b200rt_ppm_trace_photons_multi does it
 *   over peer memory inside one process)
 *   b200rt_ppm_counts_apply  n = n_at_last_sync + the summed delta, on this GPU
 * S = sum f*phi/R(k) stays local and additive: sum b200rt_ppm_export's float4 array once, when an image is wanted. */
/* Before the eye pass: this scene is share `part` of `parts` of every photon pass.  Between two synchronisations it
 * claims the acceptance indices n + part, n + part + parts, ... of a hit point (every parts-th one: the other GPUs
 * accept photons of the same stream at the same rate), so the radius r2(index) it tests against follows the global
 * count instead of lagging behind it by the other GPUs' share.  Default (1 GPU): part 0 of 1. */
int b200rt_ppm_set_share(b200rt_scene*, int part, int parts);
int b200rt_ppm_counts_delta(b200rt_scene*, uint32_t** d_delta, int64_t* n_hit_points);
int b200rt_ppm_counts_apply(b200rt_scene*);
int b200rt_ppm_eye_pass_multi(b200rt_scene* const* scenes, int n_scenes, int camera_index, uint64_t seed, B200PpmStats*);
/* One pass: opts->photons photons split by index over the scenes, traced concurrently, counts synchronised. */
int b200rt_ppm_trace_photons_multi(b200rt_scene* const* scenes, int n_scenes, const B200PpmOpts*, B200PpmStats*);
int b200rt_ppm_density_multi(b200rt_scene* const* scenes, int n_scenes, double total_num_of_photons, float* rgbw_out);

/* ---- next row: Photographic tonemap + sRGB of the frame b200rt_render* left in HBM ----
 * Replaces Photographic_tmo::apply_tmo + the gamma loop of main (HW7/src/Photographic_tmo.cpp:10-49,
 * HW7/src/main.cpp:59-91) + write_png's clamp: W*H*3 bytes into caller-owned host memory. */
int b200rt_tonemap_photographic(b200rt_scene*, int camera_index, float image_key, float burn_percent,
                                float saturation, uint8_t* rgb8_out);

/* Debug / test hook: copies `count` elements of one wavefront-pool array of the LAST render batch to
 * host memory.  field: 0 ray_o(f4) 1 ray_d(f4) 2 hit_a(f4) 3 nrm(f4) 4 rec0(f4) 5 rec1(f4) 6 rec2(f4)
 * 7 rec4(f4) 8 final_c(f4) 9 kd(f4) 10 pixel(u32) 11 key(u32) 12 hit_b(i32x2) 13 rec3(u32). */
int b200rt_debug_pool(b200rt_scene*, int field, void* out, int64_t count);

/* Parity hook (HW5..HW7 shading): with enable != 0 every later render of the scene counts, per pixel and light, the
 * shadow tests that came out "lit" (HW7/src/Scene.cpp:283-290) into a W*H*n_lights int32 array, cleared at the start
 * of each render; b200rt_debug_get_lit_counts copies it to host memory.  tests/ compare it with the oracle's counts:
 * decisions may differ only for the light SHAPES, whose own surface sits at t = dist - eps up to rounding. */
int b200rt_debug_lit_counts(b200rt_scene*, int enable);
int b200rt_debug_get_lit_counts(b200rt_scene*, int camera_index, int32_t* out);

/* Test hook: the device's leaf functions on caller-supplied rows (host pointers).  which: 0 Bounding_box::intersect
 * (in 12: min, max, o, d; out 2: decision of the exact gate, of the filtered gate), 1 BRDF::get_reflectance (in 15:
 * normal, kd, ks, w_i, w_o; out 3), 2 Texture::get_color_at (in 2: u, v; out 3), 3 Perlin_noise::get_value_at (in 3;
 * out 1).  tests/test_leaf_vectors.py runs the reference's own known-answer vectors through it. */
int b200rt_debug_leaf(int device, int which, const b200rt_texture* tex, const float* texels, int64_t n_texels,
                      const b200rt_brdf* brdf, const float* in, int rows, int n_in, float* out, int n_out);

/* The CUDA stream (cudaStream_t) every launch of this scene is issued on, so a
 * caller can bracket calls with its own events. */
void* b200rt_scene_stream(b200rt_scene*);
/* Bytes of scene data resident in HBM (nodes + triangles + tables). */
uint64_t b200rt_scene_bytes(b200rt_scene*);

/* ---- next row: binary cache of a flattened scene ----
 * The reference re-parses its XML through stringstreams on every run and knows "binaryFile" side files for vertex
 * and face data (HW7/src/Scene.cpp:998-1032, 1839-1904).  These two calls store / restore the WHOLE flattened
 * description (reference tree, SAH tree, gates, records, textures) so a host skips parse + BVH build.
 * load() returns one malloc'd block (desc + arrays) to be released with b200rt_scene_desc_free; the file carries
 * the struct sizes of this header and an FNV-1a checksum and is rejected loudly when either does not match. */
int b200rt_scene_desc_save(const B200SceneDesc*, const char* path);
int b200rt_scene_desc_load(const char* path, B200SceneDesc** out);
void b200rt_scene_desc_free(B200SceneDesc*);

/* ---- next row: the device-walked tree built on the device ----
 * Rebuilds, on the GPU, the tree k_trace_fast walks inside every mesh (LBVH: Morton order, Karras' hierarchy, bottom-up
 * boxes) from the triangles already resident, replacing the host-built SAH tree: ~ms instead of ~1 s for 1 M
 * triangles.  The reference's decisions are not in that tree (they are the triangle test, tri_gate and the tie-break
 * key), so every result stays bit-identical; traversal is slower than through the SAH tree, hence an explicit call.
 * The reference builds its own tree recursively on the host (HW7/src/Bounding_volume_hierarchy.cpp:3-29).
 * ms_out (optional): device time of the build.  b200rt_lbvh_ranges / _morton63: the hierarchy rule on the host (tests). */
int b200rt_scene_build_bvh_device(b200rt_scene*, float* ms_out);
int b200rt_lbvh_ranges(const uint64_t* sorted_keys, int n, int32_t* first_last_split);
uint64_t b200rt_lbvh_morton63(float x, float y, float z);

/* ---- next row (second half): refit for geometry that moves between frames ----
 * The reference has one answer to a moved object: parse and build everything again (its Scene is immutable;
 * per-ray motion blur is Translation(velocity * time) * M inverted inside Mesh_instance::intersect,
 * HW7/include/Mesh.h:92-104, which the device restates per ray because the 4x4 cofactor inverse of the composed matrix
 * is what decides the hit bits).  For a host that animates a scene frame by frame:
 *   b200rt_builder_set_instance_transform / _set_sphere_transform   new matrix and velocity of a top-level object
 *       (instance index: as b200rt_builder_add_instance returned it; mesh m's base instance is index m when all meshes
 *       are added first, as the loader does)
 *   b200rt_builder_refit    re-runs ONLY the top-level half of b200rt_builder_finish — inverse and normal matrices,
 *       world boxes (Mesh_instance ctor, HW3/include/Mesh.h:101-111), Scene::bvh over the objects, ranks; the result
 *       equals a fresh build of the moved scene byte for byte; mesh trees, triangle records and gates are reused
 *   b200rt_scene_refit      copies that top-level half over the resident scene: kilobytes instead of the whole
 *       description; meshes, both of their trees and all triangle data stay in HBM.  ms_out: device time of the copy.
 * Deforming meshes (vertices change) are the other half of rank 1: b200rt_scene_build_bvh_device above. */
int b200rt_builder_set_instance_transform(b200rt_builder*, int instance, const float xf[16], const float velocity[3]);
int b200rt_builder_set_sphere_transform(b200rt_builder*, int sphere, const float xf[16], const float velocity[3]);
int b200rt_builder_refit(b200rt_builder*, const B200SceneDesc** out);
int b200rt_scene_refit(b200rt_scene*, const B200SceneDesc*, float* ms_out);

/* Measurement aid: streaming-read bandwidth over a `bytes` working set, `iters` passes, in GB/s.
 * 48 MB measures the L2 (the denominator for scene fetches when the tree is L2-resident), >= 1 GB measures HBM. */
int b200rt_measure_read_bandwidth(int device, uint64_t bytes, int iters, double* gbs_out);

const char* b200rt_last_error(void);
const char* b200rt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200RT_H_ */
