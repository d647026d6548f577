/* liboisat -- C-ABI of the B200-native OI-SAT-GMI assimilation hot path.
 *
 * The reference (ahsouri/OI-SAT-GMI) is pure Python: it has no FFI layer, the
 * boundary of its hot path is a set of plain Python functions
 * (oisatgmi/interpolator.py:100, amf_recal.py:121, ak_conv_mopitt.py:8,
 * ak_conv_gosat.py:8, averaging.py:26, optimal_interpolation.py:6).  The
 * drop-in modules in oisatgmi_b200/ keep those signatures and call the entry
 * points below through ctypes; INTEGRATION.md shows the binding a maintainer
 * of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative OISAT_E_* code otherwise;
 *     oisat_last_error() gives the thread-local message of the last failure;
 *   - every array argument is a CALLER-OWNED DEVICE pointer unless the name
 *     starts with `h_` (host); extents are explicit int64; `stream` is a
 *     cudaStream_t passed as void*;
 *   - no hidden host synchronisation, no allocation in hot calls (scratch is
 *     caller-provided, sized by the *_workspace functions), no global state
 *     beyond the error string;
 *   - two-dimensional operands are "level-major": element (level l, item i)
 *     lives at base[l * stride + i].
 */
#ifndef OISAT_H_
#define OISAT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OISAT_ABI_VERSION 2

enum {
  OISAT_OK = 0,
  OISAT_E_ARG = -1,     /* bad argument */
  OISAT_E_CUDA = -2,    /* CUDA runtime/driver error */
  OISAT_E_UNSUPPORTED = -3
};

/* element types of input arrays, as delivered by the reference's readers
 * (float16 for most L2 fields, reader.py:846-883; float32 for lat/lon and the
 * model fields, reader.py:138-152) */
enum { OISAT_F16 = 1, OISAT_F32 = 2, OISAT_F64 = 3, OISAT_U8 = 4, OISAT_I32 = 5 };

/* per-field value transforms applied while gathering (interpolator.py:186:
 * sigma is squared IN THE INPUT DTYPE before the float64 mask multiply) */
enum { OISAT_OP_NONE = 0, OISAT_OP_SQUARE_NATIVE = 1 };

/* post-transform of a gridded row (interpolator.py:188: sqrt of the gridded variance) */
enum { OISAT_POST_NONE = 0, OISAT_POST_SQRT = 1 };

typedef struct oisat_field {
  const void* data;     /* device: [nlev][lev_stride] of `dtype`                 */
  int32_t dtype;        /* OISAT_F16 / F32 / F64                                  */
  int32_t op;           /* OISAT_OP_*                                             */
  int32_t post;         /* OISAT_POST_*                                           */
  int32_t nlev;         /* 1 for 2-D fields                                       */
  int64_t lev_stride;   /* elements between consecutive levels (= pixel count)   */
  double box_weight;    /* 1/(kx*ky), or 1/(kx*ky)^2 for variances (:40-46)       */
} oisat_field;

const char* oisat_last_error(void);
int oisat_abi_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t oisat_launch_count(void);

/* ---- K0: proximity predicate -------------------------------------------------
 * keep[j*W+i] = 1 iff some pixel lies within `radius` of fine-grid node
 * (xs[i], ys[j]), i.e. NOT (dists > radius) of interpolator.py:145-150,16 /
 * filler_gosat.py:132-137,17.  Euclidean distance in degrees, float64,
 * sqrt(dx*dx+dy*dy) evaluated without FMA contraction so the predicate is
 * bit-identical to scipy's cKDTree distance.  `keep` must be zeroed by the caller.
 * xs/ys must be (approximately) uniformly spaced and ascending. */
int oisat_distmask(const void* px_lon, const void* px_lat, int32_t coord_dtype, int64_t n_px,
                   const double* xs, int64_t W, const double* ys, int64_t H,
                   double radius, uint8_t* keep, void* stream);

/* ---- K1: geometry plan (triangulation on the host, the rest on the device) -----
 * HOST function: Delaunay triangulation of n points, h_tri receives 3 vertex
 * indices per triangle (capacity in triangles; 2n is always enough).  Returns the
 * triangle count, or a negative OISAT_E_* (OISAT_E_UNSUPPORTED: no triangle
 * exists -- scipy.spatial.Delaunay raises for such input and the reference
 * skips the granule, interpolator.py:152-155).  *n_ties counts exactly collinear /
 * co-circular configurations met while deciding: if it is non-zero the
 * triangulation is not unique and callers that need Qhull's choice must use it. */
int64_t oisat_h_delaunay(const double* h_x, const double* h_y, int64_t n, int32_t* h_tri,
                         int64_t tri_capacity, int64_t* n_ties);

/* HOST function: the same triangulation for the pixel centres of a swath given as
 * an n_rows x n_cols curvilinear lattice (row-major, the layout of the readers'
 * 2-D latitude_center / longitude_center that interpolator.py:131-133 flattens).
 * Points are inserted in a coarse-to-fine lattice order, each located by a short
 * walk from the previous one (*path = 1; a radial sweep needs ~10x more edge flips
 * on a long thin band).  Lattices with fewer than two lines or two points per line
 * go to the general builder above (*path = 0).  Same return value, triangle set
 * and tie report. */
int64_t oisat_h_delaunay_swath(const double* h_x, const double* h_y, int64_t n_rows,
                               int64_t n_cols, int32_t* h_tri, int64_t tri_capacity,
                               int64_t* n_ties, int32_t* path);

/* HOST function: oisat_h_delaunay_swath without its last pass.  h_half receives the twin
 * half-edge of every triangle edge (3 per triangle, -1 on the hull); *n_ties reports the
 * exact ties on the hull only, and oisat_near_ties (device) finishes the report from
 * h_tri / h_half.  *path = 0: the general builder ran instead, *n_ties is complete and
 * h_half was not written. */
int64_t oisat_h_delaunay_swath_adj(const double* h_x, const double* h_y, int64_t n_rows,
                                   int64_t n_cols, int32_t* h_tri, int64_t tri_capacity,
                                   int32_t* h_half, int64_t* n_ties, int32_t* path);

/* ---- K12: the triangulation of a swath finished on the device (round 2) -------------------
 * Replaces, for structured swaths, the host triangulation behind interpolator.py:153
 * (scipy.spatial.Delaunay inside LinearNDInterpolator).  The lattice the pixels come in is
 * already a triangulation of almost the whole footprint; the host only classifies its quads
 * and triangulates what the lattice leaves of the convex hull (outline pockets, the lips of a
 * date-line tear), the device assembles the seed and turns it into the Delaunay
 * triangulation by rounds of independent edge flips (Lawson).
 *
 * HOST: oisat_h_delaunay_seed_parts -- h_qtri[(n_rows-1)*(n_cols-1)]: first triangle of every
 *   seeded quad (-1: not seeded); h_otri / h_ohalf (3 per triangle, out_capacity triangles):
 *   the triangles outside the lattice and their twin half-edges in the numbering of the
 *   result; info[5] = {seeded quads, seam vertices, outside triangles, sigma, declining
 *   check}; *max_abs_coord (may be NULL): the largest |coordinate|.  Returns the number of triangles, 0 when the construction does not apply (the
 *   caller uses oisat_h_delaunay_swath_adj), negative OISAT_E_* on bad arguments.
 * HOST: oisat_h_delaunay_seed -- the same construction whole on the host (flags bit 0: with
 *   Lawson's flips; bit 1: with the near-tie scan), and oisat_h_flip_rounds, the device
 *   rounds replayed serially (same per-edge code, csrc/flip_rounds.h): the CPU tests' model.
 * DEVICE: oisat_seed_assemble -- tri / half (3 * (2 n_quads + n_outside) each) from the
 *   uploaded parts.  oisat_flip_delaunay -- flips in place until every edge is locally
 *   Delaunay; result[4] = {rounds, flips, edges still not Delaunay (rounds cut off), edges
 *   the floating-point filter cannot decide}: when either of the last two is non-zero the
 *   caller must take the exact host builder.  workspace: oisat_flip_workspace_bytes. */
int64_t oisat_h_delaunay_seed_parts(const double* h_x, const double* h_y, int64_t n_rows,
                                    int64_t n_cols, int32_t* h_qtri, int32_t* h_otri,
                                    int32_t* h_ohalf, int64_t out_capacity, int64_t* info,
                                    double* max_abs_coord);
int64_t oisat_h_delaunay_seed(const double* h_x, const double* h_y, int64_t n_rows,
                              int64_t n_cols, int32_t* h_tri, int64_t tri_capacity,
                              int32_t* h_half, int64_t* n_ties, int32_t flags, int64_t* info);
int oisat_h_flip_rounds(const double* h_x, const double* h_y, int32_t* h_tri, int32_t* h_half,
                        int64_t n_tri, int64_t max_rounds, int64_t* result);
int oisat_seed_assemble(const int32_t* qtri, int64_t n_rows, int64_t n_cols, int32_t sigma,
                        int64_t n_quads, const int32_t* otri, const int32_t* ohalf,
                        int64_t n_outside, int32_t* tri, int32_t* half, void* stream);
int64_t oisat_flip_workspace_bytes(int64_t n_tri);
/* several meshes (the granules of a day) through the rounds together -- a round costs two
 * barriers whatever it holds; items: HOST array; workspace: oisat_flip_workspace_bytes of the
 * total triangle count; result[4 * n_items], rounds and flips being those of the batch */
typedef struct oisat_flip_item {
  int32_t* tri;
  int32_t* half;
  int64_t n_tri;
  const void* px;
  const void* py;
} oisat_flip_item;
int oisat_flip_delaunay_batch(const oisat_flip_item* items, int32_t n_items, int32_t coord_dtype,
                              void* workspace, uint64_t* result, void* stream);
int oisat_flip_delaunay(int32_t* tri, int32_t* half, int64_t n_tri, const void* px,
                        const void* py, int32_t coord_dtype, void* workspace, uint64_t* result,
                        void* stream);

/* Device part of the tie report: edges whose fourth point lies within Qhull's tolerance of
 * the circumcircle (2e-14 x max_abs_coord^2 x twice the triangle area; exact co-circular
 * quadruples included).  tri / half: device copies of the arrays above; px / py: the
 * pixel coordinates (f32 or f64, widened exactly); *n_ties: device counter (zeroed here). */
int oisat_near_ties(const int32_t* tri, const int32_t* half, int64_t n_tri, const void* px,
                    const void* py, int32_t coord_dtype, double max_abs_coord, uint64_t* n_ties,
                    uint8_t* tri_flag, void* stream);

/* tri_flag (may be NULL; n_tri bytes, zeroed by the caller) receives 1 for both triangles of
 * every tied edge: Qhull's triangulation can differ from the exact one only inside those
 * quadrilaterals (it merges them into one facet and re-splits it its own way).
 * oisat_flagged_nodes counts the mesh nodes oisat_locate placed in a flagged triangle
 * (node_tri[f] != INT32_MAX and tri_flag[node_tri[f]]): when there is none, the stencils do
 * not depend on how the tied quadrilaterals are split and the exact triangulation serves. */
int oisat_flagged_nodes(const int32_t* node_tri, int64_t n_nodes, const uint8_t* tri_flag,
                        uint64_t* n_flagged, void* stream);

/* node_tri[f] (caller pre-fills with INT32_MAX) <- lowest index of a triangle that
 * contains mesh node f by scipy's rule (barycentric coordinates within
 * [-eps, 1+eps], eps = 100*DBL_EPSILON); only nodes with keep[f] != 0 are tested.
 * `work`: 2 * n_tri + 2 int32 of scratch. */
int oisat_locate(const int32_t* tri, int64_t n_tri, const void* px, const void* py,
                 int32_t coord_dtype, const double* xs, int64_t W, const double* ys, int64_t H,
                 const uint8_t* keep, int32_t* node_tri, int32_t* work, void* stream);

/* cell_ok[c] = nn_ok[c] and every node window[c*nwin + k] is located */
int oisat_plan_cells(const int32_t* window, int32_t nwin, const uint8_t* nn_ok, int64_t n_cell,
                     const int32_t* node_tri, uint8_t* cell_ok, void* stream);

/* stencil of the kept cells: vertices and barycentric weights of every window node,
 * pair-major ([cell][3*nwin], fused kernel) or stencil-major ([3*nwin][cell], K2) */
int oisat_plan_fill(const int32_t* cells, int64_t n_cells, const int32_t* window, int32_t nwin,
                    const int32_t* node_tri, const int32_t* tri, const void* px, const void* py,
                    int32_t coord_dtype, const double* xs, int64_t W, const double* ys,
                    int32_t pair_major, int32_t* vert, double* w, void* stream);

/* ---- nearest-neighbour gridding modes (interpolator.py:17-20 type 2, :28-33 type 4) ----
 * node_px[f] <- index of the pixel nearest to mesh node f (squared Euclidean distance
 * in degrees, float64, as scipy's KD-tree compares it; lowest index on exact ties), or
 * INT32_MAX when no pixel lies within `radius` -- those nodes are NaN in the reference
 * too (dists > 2*threshold, :19,32).  `work`: W*H uint64 of scratch. */
int oisat_nearest_pixel(const void* px_lon, const void* px_lat, int32_t coord_dtype, int64_t n_px,
                        const double* xs, int64_t W, const double* ys, int64_t H, double radius,
                        uint64_t* work, int32_t* node_px, void* stream);

/* nearest-neighbour stencil of the kept cells, in the layout of oisat_plan_fill: each
 * window node contributes the triple (pixel, pixel, pixel) with weights (1, 0, 0), so the
 * same apply / fused kernels serve both modes.  window == NULL: nwin = 1, node = cell. */
int oisat_plan_fill_nearest(const int32_t* cells, int64_t n_cells, const int32_t* window,
                            int32_t nwin, const int32_t* node_px, int32_t pair_major,
                            int32_t* vert, double* w, void* stream);

/* ---- K7: reader front-end (reader.py:807-903 OMI NO2, :906-983 OMI
 * HCHO, :707-804 TROPOMI NO2): file variables -> the arrays `interpolator` receives.
 * All bit-exact with the numpy expressions of the readers (promotion rules included).
 *
 * out[i] = float16(((src[i] * f0) * f1) * f2), products in src's dtype (float32 or
 * float64); n_factors = 0 is a plain cast (any float dtype or int32).  h_factors: HOST. */
int oisat_reader_scale_f16(const void* src, int32_t dtype, int64_t n, const double* h_factors,
                           int32_t n_factors, void* out, void* stream);

/* quality_flag (float64).  mode 0, OMI NO2 (reader.py:849-870): -100 when bits 0 and 1 of
 * the flag are both set, else 1; times float16(cloud) < float16(0.3); times
 * float16(terrain) < float16(0.2).  mode 1, OMI HCHO (:939-949): (flag == 0) *
 * (float16(cloud) < float16(0.4)); terrain unused. */
int oisat_reader_quality(int32_t mode, const void* flags, int32_t flags_dtype, const void* cloud,
                         int32_t cloud_dtype, const void* terrain, int32_t terrain_dtype, int64_t n,
                         double* quality_flag, void* stream);

/* scattering weights as float16 [n_lev][n_px]: cast, optional per-pixel factor (TROPOMI:
 * averaging kernel x total AMF, product in the factor's dtype, :773-774), then NaN / inf /
 * > 100 / < 0 -> 0 (:887-888).  pixel_major: src is [n_px][n_lev] (OMI NO2, TROPOMI files). */
int oisat_reader_weights(const void* src, int32_t dtype, int32_t pixel_major, int32_t n_lev,
                         int64_t n_px, const void* scale, int32_t scale_dtype, void* out,
                         void* stream);

/* mid-level pressures as float16 [n_lev][n_px].  mode 0: out[l][.] = a[l] (OMI NO2's fixed
 * levels).  mode 1 (OMI HCHO, :965-966): x = float16(ps); 0.5*((a[l]+b[l]x)+(a[l+1]+b[l+1]x)).
 * mode 2 (TROPOMI, :763,772-773): x = float32(ps)/float32(ps_div);
 * 0.5*(((a[l]+b[l]x)+a[l+1])+b[l+1]x) in float64; mode 3: the same expression in float32
 * (files that store the tm5 coefficients as float32).  a, b: n_lev+1 float64 on the device. */
int oisat_reader_pmid(int32_t mode, const double* a, const double* b, const void* ps,
                      int32_t ps_dtype, double ps_div, int32_t n_lev, int64_t n_px, void* out,
                      void* stream);

/* tropopause[p] = p_mid[layer[p]][p] when 0 < layer[p] < n_lev, else NaN (:781-789) */
int oisat_reader_tropopause(const int32_t* layer, const void* p_mid, int32_t n_lev, int64_t n_px,
                            void* out, void* stream);

/* MOPITT / GOSAT clean-up (reader.py:1143-1203 mopitt_reader_co, :1228-1262 gosat_reader_xch4):
 * out[lev][px] = cast_out(post(factors(pre(src)))) with pre bit 0: v <= 0 -> NaN, bit 1: inf ->
 * NaN on the source value; up to three HOST factors multiplied one after the other in the
 * source dtype (float32 / float64); post bit 0: r <= 0 -> NaN on the converted value;
 * pixel_major != 0: src is [px][lev] (the files' profile variables) and is transposed. */
int oisat_reader_clean(const void* src, int32_t dtype, int64_t n_px, int32_t n_lev,
                       int32_t pixel_major, int32_t pre, const double* h_factors,
                       int32_t n_factors, int32_t out_dtype, int32_t post, void* out, void* stream);
/* MOPITT x_col (reader.py:1168): float32(float16(1e6 * vcd16) / float32(dry * 1e-15)). */
int oisat_reader_mopitt_xcol(const void* vcd_f16, const float* dry_air, int64_t n, float* x_col,
                             void* stream);

/* good[p] = (quality_flag[p] > thresh)  (interpolator.py:126-128) */
int oisat_quality_mask(const void* qflag, int32_t dtype, int64_t n_px, double thresh,
                       uint8_t* good, void* stream);

/* ---- K2: apply a geometry plan to many fields ---------------------------------
 * The plan is a per-output-cell stencil: cell c receives, for every row r,
 *   out[r][oidx(c)] = post( sum_{k<nwin} box_weight * ( (0 + w[3k][c]*z0) + w[3k+1][c]*z1 ) + w[3k+2][c]*z2 )
 * with z = op(field value at vertex vert[3k+j][c]) or NaN when the vertex pixel
 * is not `good` -- the composition of LinearNDInterpolator (interpolator.py:13-15),
 * the box filter (:66-76) and the nearest-node sampling (:78-91), SURVEY.md App. E.
 * vert/w are [3*nwin][n_cells] (stencil-major, cell-minor).  out_index==NULL
 * writes compactly (oidx(c)=c), otherwise oidx(c)=out_index[c] (dense scatter;
 * the caller pre-fills `out` with NaN).  `h_fields` is a HOST array. */
int oisat_interp_apply(const int32_t* vert, const double* w, int32_t nwin, int64_t n_cells,
                       const uint8_t* good, const oisat_field* h_fields, int32_t n_fields,
                       double* out, int64_t out_stride, const int32_t* out_index,
                       void* stream);

/* ---- K3: per-cell vertical operators -----------------------------------------
 * Common layout: item i uses satellite column sat_index[i] (NULL: i) of the
 * level-major satellite arrays (stride sat_stride) and model column
 * ctm_index[i] (NULL: sat column) of the level-major model arrays (stride
 * ctm_stride).  Results are written at the satellite column.
 *
 * ctm_mode 0: native model fields, float32 (delta_p, profile, p_mid): the
 *             partial column, its log-pressure and the column sum are evaluated
 *             in float32 exactly as numpy does (amf_recal.py:51-56,108,116);
 * ctm_mode 1: model fields already resampled to the satellite grid and held as
 *             float64 (amf_recal.py:58-83): `ctm_a` is p_mid, `ctm_b` the
 *             partial column / profile, `ctm_c` (AK operators) the air column. */

/* AMF recalculation, amf_recal.py:93-119,175-183.  trop may be NULL. */
int oisat_vertical_amf(int64_t n_items, const int32_t* sat_index, const int32_t* ctm_index,
                       const double* vcd, const double* amf, const double* trop,
                       const double* p_sat, const double* sw, int32_t n_sat_lev, int64_t sat_stride,
                       const void* ctm_pmid, const void* ctm_b, const void* ctm_dp,
                       int32_t ctm_mode, int32_t n_ctm_lev, int64_t ctm_stride,
                       double* new_amf, double* ctm_vcd, double* vcd_out, void* stream);

/* model column without scattering weights (O3), amf_recal.py:160-171 */
int oisat_vertical_column(int64_t n_items, const int32_t* sat_index, const int32_t* ctm_index,
                          const double* vcd, const double* trop,
                          const void* ctm_pmid, const void* ctm_b, const void* ctm_dp,
                          int32_t ctm_mode, int32_t n_ctm_lev, int64_t ctm_stride,
                          double* ctm_vcd, void* stream);

/* MOPITT averaging-kernel convolution, ak_conv_mopitt.py:118-142.
 * ak has n_sat_lev+1 rows (row 0 = surface). */
int oisat_vertical_mopitt(int64_t n_items, const int32_t* sat_index, const int32_t* ctm_index,
                          const double* vcd, const double* ap_col, const double* ap_sfc,
                          const double* p_sat, const double* ak, const double* ap_prof,
                          int32_t n_sat_lev, int64_t sat_stride,
                          const void* ctm_pmid, const void* ctm_prof, const void* ctm_dp_or_air,
                          int32_t ctm_mode, int32_t n_ctm_lev, int64_t ctm_stride,
                          double* ctm_vcd, double* ctm_xcol, void* stream);

/* GOSAT averaging-kernel convolution, ak_conv_gosat.py:118-141 (keyed on x_col). */
int oisat_vertical_gosat(int64_t n_items, const int32_t* sat_index, const int32_t* ctm_index,
                         const double* x_col, const double* p_sat, const double* ak,
                         const double* ap_prof, const double* pw,
                         int32_t n_sat_lev, int64_t sat_stride,
                         const void* ctm_pmid, const void* ctm_prof,
                         int32_t ctm_mode, int32_t n_ctm_lev, int64_t ctm_stride,
                         double* ctm_xcol, void* stream);

/* ---- K6: model -> satellite-grid resampling ---------------------------------
 * out[l][i] = mean over the (ky,kx) window (symmetric edge reflection) anchored at
 * node nn[i] of level l of `src` (float32 or float64, [nlev][H*W]), NaN when
 * nn_ok[i]==0: interpolator._upscaler used with the grids' roles swapped
 * (amf_recal.py:58-83, ak_conv_mopitt.py:79-110).  Output float64.
 * src_op derives the resampled quantity on the fly from float32 model fields,
 * in float32 like numpy does: OISAT_SRC_PARTIAL_COLUMN: src = delta_p, src2 =
 * mixing ratio (amf_recal.py:51-56); OISAT_SRC_AIR_COLUMN: src = delta_p
 * (ak_conv_mopitt.py:68). */
enum { OISAT_SRC_VALUE = 0, OISAT_SRC_PARTIAL_COLUMN = 1, OISAT_SRC_AIR_COLUMN = 2 };
int oisat_grid_resample(const void* src, const void* src2, int32_t src_op, int32_t dtype,
                        int32_t nlev, int64_t H, int64_t W,
                        int32_t ky, int32_t kx, double box_weight,
                        const int32_t* nn, const uint8_t* nn_ok, int64_t n_out,
                        double* out, int64_t out_stride, void* stream);

/* ---- K4: temporal accumulation (averaging.py:64-108, 11-24) ------------------
 * acc is [10][n_cell] float64: rows 0-4 running sums of (sat vcd, sigma^2,
 * model vcd, aux1, aux2), rows 5-9 the matching counts (exact integers held as
 * float64 so one all-reduce covers the block).  Granules are added one call at
 * a time IN ORDER, which reproduces numpy's sequential axis-0 nanmean bit for
 * bit.  Any of the five inputs may be NULL (treated as all-NaN). */
int oisat_accum_add(double* acc, int64_t n_cell, const double* vcd, const double* sigma,
                    const double* ctm_vcd, const double* aux1, const double* aux2, void* stream);
/* The same with the second field already squared (sigma^2 evaluated by the caller in the
 * uncertainty array's own dtype, as numpy does for float16 / float32 grids, averaging.py:101). */
int oisat_accum_add_variance(double* acc, int64_t n_cell, const double* vcd,
                             const double* variance, const double* ctm_vcd, const double* aux1,
                             const double* aux2, void* stream);

/* means and the error sqrt(sum sigma^2 / n^2); outputs may alias nothing in acc */
int oisat_accum_finalize(const double* acc, int64_t n_cell, double* sat_vcd, double* sat_err,
                         double* ctm_vcd, double* aux1, double* aux2, void* stream);

/* ---- K5: optimal interpolation (optimal_interpolation.py:14-52, driver.py:65-114)
 * prepare: y = clip((y - bias_a)/bias_b, 0) in place (driver.py:65-106 then
 *          optimal_interpolation.py:14), Sa = (xa*err_pct/100)^2, So = sigma^2. */
int oisat_oi_prepare(const double* xa, double* y, const double* sigma, int64_t n,
                     double bias_a, double bias_b, double err_pct,
                     double* Sa, double* So, void* stream);

/* sweep: for every regularisation factor r (HOST array h_factors, n_factors<=128)
 * the sum and count of the finite entries of AK_r = 1 - Sb_r/(Sa r), summed in
 * numpy's pairwise order so that nanmean(AK) is reproduced bit for bit.
 * sums/counts: [n_factors] float64 device outputs.  `work` holds
 * oisat_oi_sweep_workspace(n, n_factors) bytes. */
int64_t oisat_oi_sweep_workspace(int64_t n, int32_t n_factors);
int oisat_oi_sweep(const double* Sa, const double* So, int64_t n,
                   const double* h_factors, int32_t n_factors,
                   double* sums, double* counts, void* work, int64_t work_bytes, void* stream);

/* apply the chosen factor: xb = xa + K (y - xa), AK, increment, sqrt(Sb) */
int oisat_oi_apply(const double* xa, const double* y, const double* Sa, const double* So,
                   int64_t n, double factor, double* xb, double* ak, double* inc, double* err,
                   void* stream);

/* The knee of the (factor, mean AK) curve on the device (Kneedle as the reference uses it,
 * optimal_interpolation.py:37-41: concave, increasing, S = 1, first knee; index 0 when there
 * is none) from the sums / counts of oisat_oi_sweep: *pick (int32) and *factor (the chosen
 * factor) are device outputs, means (may be NULL) receives the n_factors nanmeans.
 * oisat_oi_apply_dev is oisat_oi_apply with the factor read from device memory, so that
 * sweep -> knee -> update needs no host round trip. */
int oisat_oi_knee(const double* h_factors, int32_t n_factors, const double* sums,
                  const double* counts, int32_t* pick, double* factor, double* means, void* stream);
int oisat_oi_apply_dev(const double* xa, const double* y, const double* Sa, const double* So,
                       int64_t n, const double* factor, double* xb, double* ak, double* inc,
                       double* err, void* stream);

/* ---- K8: the data side of driver.write_to_nc (driver.py:156-227) ------------------
 * out: float32 [9][n] in the order the reference stores its variables: sat_averaged_vcd,
 * ctm_averaged_vcd_prior, ctm_averaged_vcd_posterior, sat_averaged_error, ak_OI, error_OI,
 * scaling_factor (= posterior / prior in float64, NaN / inf / 0 -> 1, :203-206), aux1, aux2. */
int oisat_output_fields(int64_t n, const double* sat_vcd, const double* ctm_prior,
                        const double* ctm_posterior, const double* sat_error, const double* ak,
                        const double* error_oi, const double* aux1, const double* aux2, float* out,
                        void* stream);

/* ---- fused month pipeline for float16 `satellite_amf` products ----------------
 * (OMI NO2/HCHO, TROPOMI NO2: the BASELINE configurations)
 *
 * pack: reader-layout fields ([nlev][n_px] float16, level-major) -> one
 * pixel-major record per pixel so that a vertex gather is a few 128-bit loads:
 *   rows = [SW_0..SW_{L-1}, p_0..p_{L-1}, vcd, sigma^2, tropopause?]; nchunk =
 *   ceil(nrow/8) 16-byte chunks per record; row r sits in chunk r % nchunk at
 *   element r / nchunk (DESIGN.md "packed record").
 * sigma is squared in float16 here (interpolator.py:186).  `good` folds the
 * quality flag (interpolator.py:126-128). */
int64_t oisat_pack_record_halfs(int32_t n_sat_lev, int32_t has_trop);
int oisat_pack_granule(const void* sw, const void* p_mid, int32_t n_sat_lev,
                       const void* vcd, const void* sigma, const void* trop /* may be NULL */,
                       int64_t n_px, void* records, void* stream);

/* a whole batch of granules in ONE launch, quality mask included: a pixel with
 * NOT(qflag > flag_thresh) becomes an all-NaN record and a NaN in amf_masked, which
 * is exactly what multiplying every field by the 1/NaN mask does
 * (interpolator.py:126-128,163).  `items` is a DEVICE array; block0 is the running
 * sum of oisat_pack_blocks(n_px) over the preceding items. */
typedef struct oisat_pack_item {
  const void* sw;      /* [L][n_px] float16 */
  const void* p_mid;   /* [L][n_px] float16 */
  const void* vcd;     /* [n_px] float16 */
  const void* sigma;   /* [n_px] float16 */
  const void* trop;    /* [n_px] float16 or NULL */
  const void* qflag;   /* [n_px] of qflag_dtype */
  const void* amf;     /* [n_px] of amf_dtype */
  int64_t n_px;
  int64_t px0;         /* first record / good byte of this granule */
  int64_t block0;      /* first thread block of this granule */
} oisat_pack_item;
int64_t oisat_pack_blocks(int64_t n_px);
int oisat_pack_batch(const oisat_pack_item* items, int32_t n_items, int64_t total_blocks,
                     int32_t n_sat_lev, int32_t has_trop, int32_t qflag_dtype,
                     double flag_thresh, int32_t amf_dtype, void* records, double* amf_masked,
                     void* stream);
/* The same with the granule of every block given by the caller (block_item[b] = index of
 * the item whose tile block b packs, [total_blocks] device int32; NULL = found by
 * bisection over items[].block0, ~9 dependent loads per block). */
int oisat_pack_batch_indexed(const oisat_pack_item* items, int32_t n_items, int64_t total_blocks,
                             const int32_t* block_item, int32_t n_sat_lev, int32_t has_trop,
                             int32_t qflag_dtype, double flag_thresh, int32_t amf_dtype,
                             void* records, double* amf_masked, void* stream);
/* The same, also writing the mask itself: px_bad[px0 + p] = 1 where NOT(qflag > flag_thresh)
 * (interpolator.py:126), [total pixels] device bytes, NULL = not wanted.  oisat_pair_alive
 * reads it.  skip_masked_records != 0: the all-NaN records of masked pixels are not written
 * (a fifth of the record traffic of a typical month) -- for callers that run the fused step
 * over the live pairs only, which never read them. */
int oisat_pack_batch_masked(const oisat_pack_item* items, int32_t n_items, int64_t total_blocks,
                            const int32_t* block_item, int32_t n_sat_lev, int32_t has_trop,
                            int32_t qflag_dtype, double flag_thresh, int32_t amf_dtype,
                            void* records, double* amf_masked, uint8_t* px_bad,
                            int32_t skip_masked_records, void* stream);

/* derived model fields, once per month instead of once per granule
 * (amf_recal.py:151-152): logp = float32 log(p_mid) (:108), pcol = float32 partial
 * column (:51-56).  n = n_slots*n_lev*n_cell elements, 16-byte aligned pointers. */
int oisat_ctm_prepare(const float* pmid, const float* prof, const float* dp, int64_t n,
                      float* logp, float* pcol, void* stream);

/* fused gather-interpolation + AMF recalculation over a batch of granules.
 * One "pair" = (granule, model cell) that the geometry plan marks as reachable.
 * Pairs are grouped in tiles of <=32 consecutive cells of one model row.
 * Per pair the kernel gathers the 3*nwin stencil vertices from the granule's
 * packed records, forms the gridded column in float64, reads the model column
 * of the matched time slot, evaluates amf_recal.py:93-119,175-183 and writes
 * staged[q][pair] for q = (vcd', sigma, model vcd, new amf, old amf).
 * All descriptor arrays are device arrays; see DESIGN.md for the tile table. */
typedef struct oisat_fused_args {
  /* tile table */
  int64_t n_tiles;
  const int32_t* tile_granule;   /* [n_tiles] granule id                              */
  const int32_t* tile_cell0;     /* [n_tiles] model cell index of lane 0 of the tile  */
  const int64_t* tile_pair0;     /* [n_tiles] first pair of the tile                  */
  const uint32_t* tile_mask;     /* [n_tiles] bit l set: cell0+l is a pair            */
  /* per-pair stencil, PAIR-major: entry k of pair p at [p*3*nwin + k] */
  int64_t n_pairs;
  int32_t nwin;
  const int32_t* vert;
  const double* w;
  double box_weight;             /* 1/(kx*ky)                                         */
  double box_weight_err;         /* 1/(kx*ky)^2, variance of the window mean          */
  /* per-granule tables, [n_granules] */
  int32_t n_granules;
  const int64_t* gran_record0;   /* first record (pixel) of the granule in `records`  */
  const int64_t* gran_px0;       /* first pixel of the granule in amf_masked          */
  const int32_t* gran_slot;      /* matched model time slot                           */
  /* pixel data */
  int64_t n_records;             /* pixels in `records` (n_records * chunks < 2^32)   */
  const void* records;           /* packed float16 records (oisat_pack_batch)         */
  const double* amf_masked;      /* [total px] NaN-masked AMF (oisat_pack_batch)      */
  int32_t n_sat_lev;
  int32_t has_trop;
  /* model fields, float32 [n_slots][n_ctm_lev][n_cell]; logp/pcol from
   * oisat_ctm_prepare, p_mid only read when has_trop (tropopause mask, :111-114) */
  const float* ctm_pmid;
  const float* ctm_logp;
  const float* ctm_pcol;
  int32_t n_ctm_lev;
  int64_t n_cell;
  /* output: staged[5][n_pairs] */
  double* staged;
  /* per-pair tables (split form only): granule and model cell of every pair */
  const int32_t* pair_granule;
  const int32_t* pair_cell;
  /* optional (NULL = derive from the tables above), tile form only: the same facts per
   * pair, so that a block's first loads do not wait for one another --
   * pair_record0[p] = gran_record0[pair_granule[p]] (= gran_px0[...]: one pixel, one record),
   * pair_ctm_off[p] = gran_slot[pair_granule[p]] * n_ctm_lev * n_cell + pair_cell[p] */
  const int64_t* pair_record0;
  const uint32_t* pair_ctm_off;
  /* optional, tile form only (NULL = every pair): the pairs without a masked stencil pixel,
   * alive_pairs[0 .. *n_alive), both device memory written by oisat_pair_alive, which has
   * then already staged row 4 (old AMF) of the live pairs and all five NaNs of the others;
   * the kernel covers the listed pairs and leaves the rest alone. */
  const int32_t* alive_pairs;
  const int64_t* n_alive;
} oisat_fused_args;

/* Pairs that can hold a value: a masked pixel is a NaN vertex for every field
 * (interpolator.py:126-128,163) and a NaN vertex makes the gridded cell NaN whatever its
 * weight, so a pair with one masked pixel among its 3*nwin stencil vertices is NaN in all
 * five staged rows.  One pass over the stencil tables and the mask bytes of
 * oisat_pack_batch_masked: dead pairs get their NaNs, live pairs their gridded old AMF
 * (staged row 4; same products and summation tree as the gather kernels, same bits) and a
 * place in alive_pairs ([n_pairs] int32; *n_alive = how many; neighbours stay neighbours).
 * pair_record0 may be NULL when pair_granule and gran_px0 are given. */
int oisat_pair_alive(int64_t n_pairs, int32_t nwin, const int32_t* vert, const double* w,
                     const int64_t* pair_record0, const int32_t* pair_granule,
                     const int64_t* gran_px0, const uint8_t* px_bad, const double* amf_masked,
                     double box_weight, double* staged, int32_t* alive_pairs, int64_t* n_alive,
                     void* stream);

int oisat_fused_amf(const oisat_fused_args* h_args, void* stream);

/* Split form, same results in two launches (preferred when a record has fewer
 * than 16 chunks, i.e. for every BASELINE product): a half-warp gather that writes
 * the gridded column of each pair to `rows` ([ceil(n_pairs/32)][rows_per_pair][32]
 * float64, so that ...) and a ONE-THREAD-PER-PAIR vertical operator whose loads are
 * all coalesced across the 32 pairs of a row-buffer tile and whose interp1d
 * bracket search is a merge of the two pressure-sorted level lists.  The tile
 * table of the args is not used; pair_granule / pair_cell are. */
int64_t oisat_rows_per_pair(int32_t n_sat_lev, int32_t has_trop);
int oisat_fused_amf_split(const oisat_fused_args* h_args, double* rows, void* stream);

/* Tile form, same results bit for bit in ONE launch and without the row buffer: a block
 * owns 16 consecutive pairs, gathers their gridded columns into shared memory and runs
 * the vertical operator there on all of its threads (bisection instead of the merge, so
 * no data-dependent control flow; running sums kept in registers in numpy's order).
 * Preferred whenever a record has fewer than 16 chunks and n_sat_lev <= 62.  The tile
 * table of the args is not used; pair_granule / pair_cell are. */
int oisat_fused_amf_tile(const oisat_fused_args* h_args, void* stream);

/* Pair tables of a month, on the device, from the concatenated granule plans (pairs in granule
 * order, cells ascending inside a granule).  oisat_segment_tables: for every model cell the
 * list of its pairs in GRANULE order -- seg_start [n_cell + 1], seg_pair [n_pairs] -- by a
 * counting sort whose result is deterministic (every cell sorts its own short segment by pair
 * index at the end); `work` = n_cell int32.  oisat_pair_tables: pair_record0 / pair_ctm_off of
 * oisat_fused_args from the per-granule tables (n_slots = time slots in the model block). */
int oisat_segment_tables(const int32_t* pair_cell, int64_t n_pairs, int64_t n_cell,
                         int64_t* seg_start, int64_t* seg_pair, int32_t* work, void* stream);
int oisat_pair_tables(int64_t n_pairs, const int32_t* pair_granule, const int32_t* pair_cell,
                      const int64_t* gran_px0, const int32_t* gran_slot, int32_t n_ctm_lev,
                      int64_t n_cell, int64_t n_slots, int64_t* pair_record0,
                      uint32_t* pair_ctm_off, void* stream);

/* ---- SSMIS precipitable water ------------------------------------
 * oisat_reader_ssmis: reader.py:1292-1297 -- pwv = float32(src); > 250 -> NaN; * 0.3; >= 75 or
 *   inf -> NaN; uncertainty = pwv * 0.05; all in float32.  src: uint8 / int32 / float.
 * oisat_pwv_partial: pwv_cal.py:62,68 -- delta_p * q / g / 10000 in float32, n elements.
 * oisat_pwv_column: pwv_cal.py:91-93 -- out[c] = nansum_k(partial[k][c] / 1000) accumulated in
 *   the partial's dtype (float32 / float64) in layer order, NaN where sat_vcd[c] is NaN or inf. */
int oisat_reader_ssmis(const void* src, int32_t dtype, int64_t n, float* pwv, float* uncertainty,
                       void* stream);
int oisat_pwv_partial(const float* delta_p, const float* profile, int64_t n, float* out,
                      void* stream);
int oisat_pwv_column(const void* partial, int32_t dtype, int32_t n_lev, int64_t n_cell,
                     const double* sat_vcd, double* out, void* stream);

/* ordered segmented reduction of the staged pair values into the accumulators:
 * for model cell c the pairs seg_pair[seg_start[c] .. seg_start[c+1]) are listed
 * in granule order, so the running sums equal numpy's sequential nanmean
 * bit for bit (no floating-point atomics).  acc layout as oisat_accum_add. */
int oisat_accum_pairs(double* acc, int64_t n_cell, const int64_t* seg_start,
                      const int64_t* seg_pair, const double* staged, int64_t n_pairs,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OISAT_H_ */
