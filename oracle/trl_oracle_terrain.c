/*
 * trl_oracle_terrain.c -- CPU ORACLE, terrain generation and scrolling (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the reference's procedural terrain:
 *   cRand                         util/Rand.cpp:6-139 on std::default_random_engine
 *   cTerrainGen2D::Build* / Add*  sim/TerrainGen2D.cpp:127-653
 *   cGroundVar2D                  sim/GroundVar2D.cpp:33-96,261-450 (InitSegments / BuildSegment / Update / AddPadding)
 *   cGroundVar2D::tSegment::Init  sim/GroundVar2D.cpp:457-520
 *   cTerrainGen3D::BuildFlat      sim/TerrainGen3D.cpp:63-78,103-111,182-203
 *   cGroundVar3D                  sim/GroundVar3D.cpp:24-141,350-420 (InitSlabs / BuildSlab / Update), tSlab::Init :511-550
 *
 * Third-party arithmetic restated from its published algorithm (libstdc++, the reference's Linux toolchain):
 *   std::default_random_engine = minstd_rand0 (x <- 16807 x mod 2^31-1), std::generate_canonical<double,53> with two
 *   draws, std::uniform_int_distribution<int> by the classic down / up-scaling rejection loops; Bullet's float world
 *   transform and btHeightfieldTerrainShape::getAabb (half extents x local scaling around the origin, margin 0).
 * Pinned against the reference's own translation units: tests/test_ref_pin.py (oracle/_ref) and tests/golden/.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "trl_oracle.h"

/* ------------------------------------------------------------------------------------------
 * cRand
 * ---------------------------------------------------------------------------------------- */
#define MINSTD_M 2147483647ULL
#define MINSTD_A 16807ULL

void trlo_rand_seed(trlo_rand* r, unsigned long seed)
{ /* linear_congruential_engine::seed: c = 0, so a seed that is 0 mod m becomes 1 */
    unsigned long long s = (unsigned long long)seed % MINSTD_M;
    r->x = (s == 0) ? 1 : s;
}

static unsigned long long minstd_next(trlo_rand* r)
{
    r->x = (MINSTD_A * r->x) % MINSTD_M;
    return r->x;
}

double trlo_rand_double(trlo_rand* r)
{ /* uniform_real_distribution<double>(0, 1): generate_canonical<double, 53>, range R = m - 1, two draws */
    const double R = 2147483646.0;
    double sum = 0, tmp = 1;
    for (int k = 0; k < 2; ++k) {
        sum += (double)(minstd_next(r) - 1ULL) * tmp;
        tmp *= R;
    }
    double ret = sum / tmp;
    if (ret >= 1.0) ret = nextafter(1.0, 0.0);
    return ret * (1.0 - 0.0) + 0.0;
}

double trlo_rand_double_range(trlo_rand* r, double mn, double mx)
{ /* Rand.cpp:29-40 */
    if (mn == mx) return mn;
    double x = trlo_rand_double(r);
    return mn + (x * (mx - mn));
}

/* uniform_int_distribution on [0, urange] from a generator with range [1, m-1] (urngrange = m - 2) */
static unsigned long long uniform_uint(trlo_rand* r, unsigned long long urange)
{
    const unsigned long long urngrange = MINSTD_M - 2ULL;
    unsigned long long ret;
    if (urngrange > urange) { /* downscaling */
        const unsigned long long uerange = urange + 1;
        const unsigned long long scaling = urngrange / uerange;
        const unsigned long long past = uerange * scaling;
        do ret = minstd_next(r) - 1ULL; while (ret >= past);
        ret /= scaling;
    } else if (urngrange < urange) { /* upscaling */
        unsigned long long tmp;
        do {
            const unsigned long long uerngrange = urngrange + 1;
            tmp = uerngrange * uniform_uint(r, urange / uerngrange);
            ret = tmp + (minstd_next(r) - 1ULL);
        } while (ret > urange || ret < tmp);
    } else {
        ret = minstd_next(r) - 1ULL;
    }
    return ret;
}

int trlo_rand_int(trlo_rand* r)
{ /* uniform_int_distribution<int>(INT_MIN + 1, INT_MAX), Rand.cpp:11,57-60 */
    unsigned long long u = uniform_uint(r, 4294967294ULL);
    long long v = (long long)u + (-2147483647LL);
    return (int)v;
}

int trlo_rand_int_range(trlo_rand* r, int mn, int mx)
{ /* Rand.cpp:62-75 */
    if (mn == mx) return mn;
    int delta = mx - mn;
    int v = abs(trlo_rand_int(r));
    return mn + v % delta;
}

int trlo_rand_flip_coin(trlo_rand* r, double p) { return trlo_rand_double_range(r, 0, 1) < p; } /* Rand.cpp:136-139 */
static int rand_sign(trlo_rand* r) { return trlo_rand_flip_coin(r, 0.5) ? -1 : 1; }                /* Rand.cpp:131-134 */

/* ------------------------------------------------------------------------------------------
 * cTerrainGen2D
 * ---------------------------------------------------------------------------------------- */
static const float kVertSpacing2D = 0.025f; /* TerrainGen2D.cpp:4-10 (SHARP_TERRAIN) */

/* parameter indices, TerrainGen2D.h:13-66 */
enum { P_GAP_SPACING_MIN, P_GAP_SPACING_MAX, P_GAP_W_MIN, P_GAP_W_MAX, P_GAP_D_MIN, P_GAP_D_MAX,
       P_WALL_SPACING_MIN, P_WALL_SPACING_MAX, P_WALL_W_MIN, P_WALL_W_MAX, P_WALL_H_MIN, P_WALL_H_MAX,
       P_STEP_SPACING_MIN, P_STEP_SPACING_MAX, P_STEP_H0_MIN, P_STEP_H0_MAX, P_STEP_H1_MIN, P_STEP_H1_MAX,
       P_BUMP_H_MIN, P_BUMP_H_MAX,
       P_NGAP_SPACING_MIN, P_NGAP_SPACING_MAX, P_NGAP_DIST_MIN, P_NGAP_DIST_MAX, P_NGAP_W_MIN, P_NGAP_W_MAX,
       P_NGAP_D_MIN, P_NGAP_D_MAX, P_NGAP_COUNT_MIN, P_NGAP_COUNT_MAX,
       P_CLIFF_SPACING_MIN, P_CLIFF_SPACING_MAX, P_CLIFF_H0_MIN, P_CLIFF_H0_MAX, P_CLIFF_H1_MIN, P_CLIFF_H1_MAX,
       P_CLIFF_MINI_COUNT_MAX, P_SLOPE_DELTA_RANGE, P_SLOPE_DELTA_MIN, P_SLOPE_DELTA_MAX, P_MAX };

static const double kDefaultParams2D[P_MAX] = { /* TerrainGen2D.cpp:12-63 */
    4, 7, 0.5, 2, -2, -2, 6, 8, 0.2, 0.2, 0.25, 0.5, 5, 7, 0.1, 0.4, -0.4, -0.1, 0, 0.03,
    3, 6, 0.1, 0.4, 0.15, 0.5, -2, -2, 1, 4, 5, 7, 0.1, 0.4, -0.4, -0.1, 0, 0.25, -0.35, 0.35 };

int trlo_terrain2d_num_params(void) { return P_MAX; }
void trlo_terrain2d_default_params(double* out) { memcpy(out, kDefaultParams2D, sizeof(kDefaultParams2D)); }

typedef struct fvec { float* d; int n, cap; } fvec;
static void push(fvec* v, float x)
{
    if (v->n >= v->cap) abort(); /* the caller sizes the buffer; never reached with sane parameters */
    v->d[v->n++] = x;
}

static int calc_num_verts(double w) { return (int)ceil(w / kVertSpacing2D) + 1; } /* :464-468 */

static double add_flat(double width, fvec* out) /* :470-498 */
{
    int num_verts = calc_num_verts(width);
    int size0 = out->n, start_empty = (size0 == 0);
    float base_h = 0;
    if (!start_empty) { --num_verts; base_h = out->d[size0 - 1]; }
    for (int i = 0; i < num_verts; ++i) push(out, base_h);
    int added = out->n - size0;
    if (start_empty) --added;
    return added * kVertSpacing2D;
}

static double add_box(double spacing, double width, double depth, fvec* out) /* :500-539 */
{
    int num_verts = calc_num_verts(spacing);
    int size0 = out->n, start_empty = (size0 == 0);
    float base_h = 0;
    if (!start_empty) { --num_verts; base_h = out->d[size0 - 1]; }
    for (int i = 0; i < num_verts; ++i) push(out, base_h);
    num_verts = calc_num_verts(width) - 1;
    float gap_h = (float)(base_h + depth);
    for (int i = 0; i < num_verts; ++i) push(out, gap_h);
    push(out, base_h);
    int added = out->n - size0;
    if (start_empty) --added;
    return added * kVertSpacing2D;
}

static double add_step(double width, double height, fvec* out) /* :541-571 */
{
    int num_verts = calc_num_verts(width);
    int size0 = out->n, start_empty = (size0 == 0);
    float base_h = 0;
    if (!start_empty) { --num_verts; base_h = out->d[size0 - 1]; }
    for (int i = 0; i < num_verts; ++i) push(out, base_h);
    push(out, (float)(base_h + height));
    int added = out->n - size0;
    if (start_empty) --added;
    return added * kVertSpacing2D;
}

static void overlay_slopes(double delta_range, double delta_min, double delta_max, double init_slope, int beg, int end,
                           trlo_rand* rand, fvec* out) /* :609-641 */
{
    double curr_slope = init_slope, curr_delta_h = 0;
    double delta_mean = 0.5 * (delta_min + delta_max);
    double delta_diff = 0.5 * (delta_max - delta_min);
    for (int i = beg; i < end; ++i) {
        if (delta_min != delta_max) {
            double delta = trlo_rand_double_range(rand, 0, delta_range);
            double sign_rand = trlo_rand_double_range(rand, -1, 1);
            double sign_threshold = (curr_slope - delta_mean) / delta_diff;
            int neg = sign_rand < sign_threshold;
            delta = neg ? -delta : delta;
            curr_slope += delta;
            curr_delta_h += curr_slope * kVertSpacing2D;
        } else {
            curr_delta_h = 0;
        }
        out->d[i] += (float)curr_delta_h;
    }
}

static void overlay_bumps(double min_dh, double max_dh, int beg, int end, trlo_rand* rand, fvec* out) /* :643-653 */
{
    for (int i = beg; i < end - 1; ++i) {
        int sgn = rand_sign(rand);
        double delta = sgn * trlo_rand_double_range(rand, min_dh, max_dh);
        out->d[i] += (float)delta;
    }
}

static double build_2d(int type, double width, const double* p, trlo_rand* rand, fvec* out);

static double build_gaps(double width, const double* p, trlo_rand* rand, fvec* out) /* :132-153 */
{
    double total_w = 0;
    while (total_w < width) {
        double spacing = trlo_rand_double_range(rand, p[P_GAP_SPACING_MIN], p[P_GAP_SPACING_MAX]);
        double w = trlo_rand_double_range(rand, p[P_GAP_W_MIN], p[P_GAP_W_MAX]);
        double d = trlo_rand_double_range(rand, p[P_GAP_D_MIN], p[P_GAP_D_MAX]);
        total_w += add_box(spacing, w, d, out);
    }
    return total_w;
}

static void pick_range(double h0_min, double h0_max, double h1_min, double h1_max, trlo_rand* rand, double* min_h, double* max_h)
{ /* :166-186, :404-424 */
    int valid_h0 = (h0_min != 0 || h0_max != 0);
    int valid_h1 = (h1_min != 0 || h1_max != 0);
    if (valid_h0 && valid_h1) {
        int heads = trlo_rand_flip_coin(rand, 0.5);
        *min_h = heads ? h0_min : h1_min;
        *max_h = heads ? h0_max : h1_max;
    } else if (valid_h0) {
        *min_h = h0_min; *max_h = h0_max;
    } else {
        *min_h = h1_min; *max_h = h1_max;
    }
}

static double build_steps(double width, const double* p, trlo_rand* rand, fvec* out) /* :155-197 */
{
    double total_w = 0;
    while (total_w < width) {
        double min_h = 0, max_h = 0;
        pick_range(p[P_STEP_H0_MIN], p[P_STEP_H0_MAX], p[P_STEP_H1_MIN], p[P_STEP_H1_MAX], rand, &min_h, &max_h);
        double w = trlo_rand_double_range(rand, p[P_STEP_SPACING_MIN], p[P_STEP_SPACING_MAX]);
        double h = trlo_rand_double_range(rand, min_h, max_h);
        total_w += add_step(w, h, out);
    }
    return total_w;
}

static double build_walls(double width, const double* p, trlo_rand* rand, fvec* out) /* :199-220 */
{
    double total_w = 0;
    while (total_w < width) {
        double spacing = trlo_rand_double_range(rand, p[P_WALL_SPACING_MIN], p[P_WALL_SPACING_MAX]);
        double w = trlo_rand_double_range(rand, p[P_WALL_W_MIN], p[P_WALL_W_MAX]);
        double h = trlo_rand_double_range(rand, p[P_WALL_H_MIN], p[P_WALL_H_MAX]);
        total_w += add_box(spacing, w, h, out);
    }
    return total_w;
}

static double build_mixed(double width, const double* p, trlo_rand* rand, fvec* out) /* :236-264 */
{
    double total_w = 0;
    const double dummy_w = kVertSpacing2D;
    while (total_w < width) {
        double curr_w = 0;
        int rand_type = trlo_rand_int_range(rand, 0, 3);
        if (rand_type == 0) curr_w = build_gaps(dummy_w, p, rand, out);
        else if (rand_type == 1) curr_w = build_steps(dummy_w, p, rand, out);
        else if (rand_type == 2) curr_w = build_walls(dummy_w, p, rand, out);
        total_w += curr_w;
    }
    return total_w;
}

static double build_narrow_gaps(double width, const double* p, trlo_rand* rand, fvec* out) /* :266-298 */
{
    int count_min = (int)p[P_NGAP_COUNT_MIN] > 1 ? (int)p[P_NGAP_COUNT_MIN] : 1;
    int count_max = (int)p[P_NGAP_COUNT_MAX] > 1 ? (int)p[P_NGAP_COUNT_MAX] : 1;
    double total_w = 0;
    while (total_w < width) {
        double spacing = trlo_rand_double_range(rand, p[P_NGAP_SPACING_MIN], p[P_NGAP_SPACING_MAX]);
        int count = trlo_rand_int_range(rand, count_min, count_max + 1);
        for (int i = 0; i < count; ++i) {
            double w = trlo_rand_double_range(rand, p[P_NGAP_W_MIN], p[P_NGAP_W_MAX]);
            double d = trlo_rand_double_range(rand, p[P_NGAP_D_MIN], p[P_NGAP_D_MAX]);
            total_w += add_box(spacing, w, d, out);
            spacing = trlo_rand_double_range(rand, p[P_NGAP_DIST_MIN], p[P_NGAP_DIST_MAX]);
        }
    }
    return total_w;
}

static double build_cliffs(double width, const double* p, trlo_rand* rand, fvec* out) /* :392-462 */
{
    int mini_count_max = (int)p[P_CLIFF_MINI_COUNT_MAX];
    double total_w = 0;
    while (total_w < width) {
        double min_h = 0, max_h = 0;
        pick_range(p[P_CLIFF_H0_MIN], p[P_CLIFF_H0_MAX], p[P_CLIFF_H1_MIN], p[P_CLIFF_H1_MAX], rand, &min_h, &max_h);
        double w = trlo_rand_double_range(rand, p[P_CLIFF_SPACING_MIN], p[P_CLIFF_SPACING_MAX]);
        double h = trlo_rand_double_range(rand, min_h, max_h);
        double curr_w = 0, curr_delta_h = 0;
        int num_mini = trlo_rand_int_range(rand, 0, mini_count_max + 1);
        for (int i = 0; i < num_mini + 1; ++i) {
            const double mini_w = (i == 0) ? w : 0.1;
            double mini_h = trlo_rand_double_range(rand, curr_delta_h, h);
            mini_h = (i == num_mini) ? h : mini_h;
            double dh = mini_h - curr_delta_h;
            curr_w += add_step(mini_w, dh, out);
            curr_delta_h = mini_h;
        }
        total_w += curr_w;
    }
    return total_w;
}

/* terrain types: cGround::eType, sim/Ground.h:26-55 */
enum { T_PLANE, T_FLAT, T_GAPS, T_STEPS, T_WALLS, T_BUMPS, T_MIXED, T_NARROW_GAPS, T_SLOPES, T_SLOPES_GAPS,
       T_SLOPES_WALLS, T_SLOPES_STEPS, T_SLOPES_MIXED, T_SLOPES_NARROW_GAPS, T_CLIFFS, T_VAR3D_FLAT };

static double build_2d(int type, double width, const double* p, trlo_rand* rand, fvec* out)
{ /* GetTerrainFunc :86-123 and the Build* wrappers :127-131,222-234,300-390 */
    int base = -1;
    switch (type) {
    case T_FLAT: return add_flat(width, out);
    case T_GAPS: return build_gaps(width, p, rand, out);
    case T_STEPS: return build_steps(width, p, rand, out);
    case T_WALLS: return build_walls(width, p, rand, out);
    case T_MIXED: return build_mixed(width, p, rand, out);
    case T_NARROW_GAPS: return build_narrow_gaps(width, p, rand, out);
    case T_CLIFFS: return build_cliffs(width, p, rand, out);
    case T_BUMPS: {
        int beg = out->n;
        double total_w = add_flat(width, out);
        overlay_bumps(p[P_BUMP_H_MIN], p[P_BUMP_H_MAX], beg, out->n, rand, out);
        return total_w;
    }
    case T_SLOPES: base = T_FLAT; break;
    case T_SLOPES_GAPS: base = T_GAPS; break;
    case T_SLOPES_STEPS: base = T_STEPS; break;
    case T_SLOPES_WALLS: base = T_WALLS; break;
    case T_SLOPES_MIXED: base = T_MIXED; break;
    case T_SLOPES_NARROW_GAPS: base = T_NARROW_GAPS; break;
    default: return add_flat(width, out);
    }
    int beg = out->n;
    double total_w = build_2d(base, width, p, rand, out);
    overlay_slopes(fabs(p[P_SLOPE_DELTA_RANGE]), p[P_SLOPE_DELTA_MIN], p[P_SLOPE_DELTA_MAX], 0, beg, out->n, rand, out);
    return total_w;
}

double trlo_terrain_gen_2d(int type, double width, const double* params, trlo_rand* rand, float* data, int* inout_count,
                           int cap)
{
    fvec v = { data, *inout_count, cap };
    double w = build_2d(type, width, params ? params : kDefaultParams2D, rand, &v);
    *inout_count = v.n;
    return w;
}

/* ------------------------------------------------------------------------------------------
 * Bullet-side float round trips shared by the 2-D segments and 3-D slabs
 * ---------------------------------------------------------------------------------------- */
static void piece_place(trlo_terrain_piece* pc, const double origin[3], double x_scale, double z_scale, double world_scale,
                        int res_x, int res_z)
{
    /* cWorld::SetPos (sim/World.cpp:343-362): float(scale) * btVector3(float(pos)) ; GetPos (:333-341): / scale */
    float sc = (float)world_scale;
    for (int k = 0; k < 3; ++k) {
        float o = sc * (float)origin[k];
        pc->origin[k] = (double)o / world_scale;
    }
    /* setLocalScaling(btVector3(float(x_scale), 1, float(z_scale))) ; GetScaling = bt_scale / world_scale */
    float ls[3] = { (float)x_scale, 1.0f, (float)z_scale };
    for (int k = 0; k < 3; ++k) pc->scale[k] = (double)ls[k] / world_scale;
    /* btHeightfieldTerrainShape::getAabb in x / z (float): center -+ (res - 1) * local_scaling * 0.5, then / scale */
    float hx = ((float)(res_x - 1) - 0.0f) * ls[0] * 0.5f;
    float hz = ((float)(res_z - 1) - 0.0f) * ls[2] * 0.5f;
    float cx = sc * (float)origin[0], cz = sc * (float)origin[2];
    pc->aabb[0] = (double)(cx - hx) / world_scale;
    pc->aabb[1] = (double)(cx + hx) / world_scale;
    pc->aabb[2] = (double)(cz - hz) / world_scale;
    pc->aabb[3] = (double)(cz + hz) / world_scale;
    pc->res[0] = res_x;
    pc->res[1] = res_z;
}

/* ------------------------------------------------------------------------------------------
 * cGroundVar2D
 * ---------------------------------------------------------------------------------------- */
#define SEG_ROWS 3 /* tSegment::gGridLength */
enum { ALIGN_MIN = 0, ALIGN_MAX = 1 };

static int seg_id(const trlo_ground_var2d* g, int s) { return g->flip_seg ? (s == 0 ? 1 : 0) : s; } /* :316-323 */
static int seg_align(const trlo_ground_var2d* g, int s)                                              /* :349-356 */
{
    if (g->flip_seg) return (s == 0) ? ALIGN_MIN : ALIGN_MAX;
    return (s == 0) ? ALIGN_MAX : ALIGN_MIN;
}
static int seg_empty(const trlo_ground_var2d* g, int id) { return g->piece[id].count == 0; }
static double seg_min_x(const trlo_ground_var2d* g, int id) { return seg_empty(g, id) ? INFINITY : g->piece[id].aabb[0]; }
static double seg_max_x(const trlo_ground_var2d* g, int id) { return seg_empty(g, id) ? -INFINITY : g->piece[id].aabb[1]; }

static void segment_init(trlo_ground_var2d* g, int id, int grid_width, double min_x)
{ /* tSegment::Init :457-520 */
    trlo_terrain_piece* pc = &g->piece[id];
    float* d = pc->data;
    double world_scale = g->world_scale;
    double aabb_min[3] = { min_x, INFINITY, 0 };
    double aabb_max[3] = { min_x + (grid_width - 1) * (double)kVertSpacing2D, -INFINITY, 0 };
    for (int i = 0; i < grid_width; ++i) {
        double h = d[i];
        if (h < aabb_min[1]) aabb_min[1] = h;
        if (h > aabb_max[1]) aabb_max[1] = h;
        d[i] *= (float)world_scale;
    }
    for (int r = 1; r < SEG_ROWS; ++r) memcpy(d + (size_t)r * grid_width, d, sizeof(float) * (size_t)grid_width);
    pc->count = grid_width * SEG_ROWS;
    double origin[3];
    for (int k = 0; k < 3; ++k) origin[k] = 0.5 * (aabb_min[k] + aabb_max[k]);
    piece_place(pc, origin, (double)kVertSpacing2D * world_scale, 1.0 * world_scale, world_scale, grid_width, SEG_ROWS);
}

static void build_segment(trlo_ground_var2d* g, int id, double bound_min, double bound_max, int align, const double fix_point[2])
{ /* BuildSegment :358-388 with AddPadding :390-397 */
    trlo_terrain_piece* pc = &g->piece[id];
    fvec v = { pc->data, 0, TRLO_TERRAIN_PIECE_CAP / SEG_ROWS };
    pc->count = 0;
    int contains_origin = (bound_min <= 0) && (bound_max >= 0);
    if (contains_origin) {
        double flat_w = fmin(bound_max - bound_min, 1 - bound_min);
        add_flat(flat_w, &v);
    }
    build_2d(g->type, bound_max - bound_min, g->params, &g->rand, &v);
    int num_verts = v.n;
    float end_h = 0;
    if (num_verts > 0) end_h = (align == ALIGN_MIN) ? v.d[0] : v.d[num_verts - 1];
    float h_offset = (float)(fix_point[1] - end_h);
    for (int i = 0; i < num_verts; ++i) v.d[i] += h_offset;
    double new_bound_min = (align == ALIGN_MIN) ? bound_min : (bound_max - (num_verts - 1) * (double)kVertSpacing2D);
    segment_init(g, id, num_verts, new_bound_min);
}

static void init_segments(trlo_ground_var2d* g, double bound_min_x, double bound_max_x)
{ /* InitSegments :290-307 */
    g->piece[0].count = g->piece[1].count = 0; /* ClearSegments :309-314 */
    g->flip_seg = 0;
    double mid = 0.5 * (bound_max_x + bound_min_x);
    double w = g->ground_width;
    for (int i = 0; i < 2; ++i) {
        int id = seg_id(g, i);
        double fix_point[2] = { 0, 0 };
        int align = seg_align(g, i);
        double bmin = (align == ALIGN_MAX) ? -w : 0;
        double bmax = (align == ALIGN_MAX) ? 0 : w;
        bmin += mid;
        bmax += mid;
        build_segment(g, id, bmin, bmax, align, fix_point);
    }
}

void trlo_ground2d_init(trlo_ground_var2d* g, int type, const double* params, double ground_width, double world_scale,
                        unsigned long seed, double bound_min_x, double bound_max_x)
{ /* cGroundVar2D::Init :33-46 (+ cGround::SetupRandGen, sim/Ground.cpp:247-253) */
    memset(g, 0, sizeof(*g));
    g->type = type;
    memcpy(g->params, params ? params : kDefaultParams2D, sizeof(kDefaultParams2D));
    g->ground_width = ground_width;
    g->world_scale = world_scale;
    trlo_rand_seed(&g->rand, seed);
    init_segments(g, bound_min_x, bound_max_x);
}

void trlo_ground2d_update(trlo_ground_var2d* g, double bound_min_x, double bound_max_x)
{ /* cGroundVar2D::Update :48-96 */
    int min_id = seg_id(g, 0), max_id = seg_id(g, 1);
    double min_x = seg_min_x(g, min_id), max_x = seg_max_x(g, max_id);
    if (bound_max_x < max_x && bound_min_x > min_x) {
        /* nothing to do */
    } else if (bound_max_x <= min_x || bound_min_x >= max_x) {
        init_segments(g, bound_min_x, bound_max_x);
    } else {
        int id, align;
        double seg_min, seg_max, fix_point[2] = { 0, 0 };
        if (bound_max_x >= max_x) {
            id = seg_id(g, 0);
            seg_min = max_x;
            seg_max = seg_min + g->ground_width;
            const trlo_terrain_piece* mx = &g->piece[max_id];
            fix_point[0] = max_x;
            fix_point[1] = mx->data[mx->count - 1] / g->world_scale; /* GetEndHeight :646-650 */
            align = ALIGN_MIN;
        } else {
            id = seg_id(g, 1);
            seg_max = min_x;
            seg_min = seg_max - g->ground_width;
            const trlo_terrain_piece* mn = &g->piece[min_id];
            fix_point[0] = min_x;
            fix_point[1] = mn->data[0] / g->world_scale; /* GetStartHeight :640-644 */
            align = ALIGN_MAX;
        }
        build_segment(g, id, seg_min, seg_max, align, fix_point);
        g->flip_seg = (seg_id(g, 0) == 0);
    }
}

/* the pieces in the reference's scan order (GetSegment(s)), as the sampling oracle consumes them */
void trlo_ground2d_view(const trlo_ground_var2d* g, trlo_ground* out)
{
    memset(out, 0, sizeof(*out));
    out->kind = TRLO_GROUND_VAR2D;
    out->num_pieces = 2;
    for (int s = 0; s < 2; ++s) {
        const trlo_terrain_piece* pc = &g->piece[seg_id(g, s)];
        out->data[s] = pc->data;
        out->res[s][0] = pc->res[0];
        out->res[s][1] = pc->res[1];
        for (int k = 0; k < 3; ++k) { out->origin[s][k] = pc->origin[k]; out->scale[s][k] = pc->scale[k]; }
        /* ContainsPos :697-700: GetMinX / GetMaxX = the rigid body's AABB (:540-564); z is not tested */
        out->bounds[s][0] = pc->aabb[0];
        out->bounds[s][1] = pc->aabb[1];
        out->bounds[s][2] = -INFINITY;
        out->bounds[s][3] = INFINITY;
    }
}

/* ------------------------------------------------------------------------------------------
 * cGroundVar3D with cTerrainGen3D::BuildFlat
 * ---------------------------------------------------------------------------------------- */
static int calc_res_3d(double size, double spacing) { return (int)ceil(size / spacing - 0.000001) + 1; } /* TerrainGen3D.cpp:70-78 */

static void build_slab(trlo_ground_var3d* g, int s, const double bmin[3], const double bmax[3])
{ /* BuildSlab :389-414, BuildFlat / AddFlat (TerrainGen3D.cpp:103-111,182-203), tSlab::Init :511-550 */
    trlo_terrain_piece* pc = &g->piece[g->slab_order[s]];
    double size[3] = { bmax[0] - bmin[0], bmax[1] - bmin[1], bmax[2] - bmin[2] };
    int res_x = calc_res_3d(size[0], g->spacing_x), res_z = calc_res_3d(size[2], g->spacing_z);
    if ((long)res_x * res_z > TRLO_TERRAIN_PIECE_CAP) abort();
    pc->count = res_x * res_z;
    for (int i = 0; i < pc->count; ++i) pc->data[i] = (float)(-INFINITY); /* ClearSlabData :441-445 */
    for (int j = 0; j < res_z; ++j)
        for (int i = 0; i < res_x; ++i) pc->data[(size_t)j * res_x + i] = (float)0.0;
    double mn[3] = { bmin[0], INFINITY, bmin[2] }, mx[3] = { bmax[0], -INFINITY, bmax[2] };
    /* tSlab::Clear leaves mMin / mMax = +-inf; BuildSlab then sets them to the bounds (y = 0 of the bound vectors)  */
    mn[1] = bmin[1];
    mx[1] = bmax[1];
    for (int i = 0; i < pc->count; ++i) {
        double h = pc->data[i];
        if (h < mn[1]) mn[1] = h;
        if (h > mx[1]) mx[1] = h;
        pc->data[i] *= (float)g->world_scale;
    }
    for (int k = 0; k < 3; ++k) { pc->min[k] = mn[k]; pc->max[k] = mx[k]; }
    double origin[3];
    for (int k = 0; k < 3; ++k) origin[k] = 0.5 * (mx[k] + mn[k]);
    piece_place(pc, origin, g->spacing_x * g->world_scale, g->spacing_z * g->world_scale, g->world_scale, res_x, res_z);
}

static void slab_bounds(int s, double w, const double mid[3], double bmin[3], double bmax[3])
{ /* :358-369 / :120-129 */
    bmin[0] = (s % 2 == 0) ? -w : 0;
    bmax[0] = (s % 2 == 0) ? 0 : w;
    bmin[1] = bmax[1] = 0;
    bmin[2] = (s < 2) ? -w : 0;
    bmax[2] = (s < 2) ? 0 : w;
    for (int k = 0; k < 3; ++k) { bmin[k] += mid[k]; bmax[k] += mid[k]; }
}

static void init_slabs(trlo_ground_var3d* g, const double bound_min[3], const double bound_max[3])
{ /* InitSlabs :350-372 */
    double mid[3];
    for (int k = 0; k < 3; ++k) mid[k] = 0.5 * (bound_max[k] + bound_min[k]);
    for (int s = 0; s < 4; ++s) {
        double bmin[3], bmax[3];
        slab_bounds(s, g->ground_width, mid, bmin, bmax);
        build_slab(g, s, bmin, bmax);
    }
}

void trlo_ground3d_init(trlo_ground_var3d* g, double spacing_x, double spacing_z, double world_scale,
                        const double bound_min[3], const double bound_max[3])
{ /* cGroundVar3D::Init :24-36 (mGroundWidth forced to 40), ResetParams: slab order 0..3 */
    memset(g, 0, sizeof(*g));
    g->ground_width = 40;
    g->spacing_x = spacing_x;
    g->spacing_z = spacing_z;
    g->world_scale = world_scale;
    for (int s = 0; s < 4; ++s) g->slab_order[s] = s;
    init_slabs(g, bound_min, bound_max);
}

void trlo_ground3d_update(trlo_ground_var3d* g, const double bound_min[3], const double bound_max[3])
{ /* cGroundVar3D::Update :38-141 */
    const trlo_terrain_piece* sl[4];
    for (int s = 0; s < 4; ++s) sl[s] = &g->piece[g->slab_order[s]];
    double curr_min[3] = { fmax(sl[0]->min[0], sl[2]->min[0]), 0, fmax(sl[0]->min[2], sl[1]->min[2]) }; /* GetBounds :422-438 */
    double curr_max[3] = { fmin(sl[1]->max[0], sl[3]->max[0]), 0, fmin(sl[2]->max[2], sl[3]->max[2]) };
    if (bound_min[0] > curr_min[0] && bound_min[2] > curr_min[2] && bound_max[0] < curr_max[0] && bound_max[2] < curr_max[2]) {
        return;
    } else if ((bound_min[0] > curr_max[0] && bound_min[2] > curr_max[2]) ||
               (bound_max[0] < curr_min[0] && bound_max[2] < curr_min[2])) {
        init_slabs(g, bound_min, bound_max);
        return;
    }
    int need_update[4] = { 0, 0, 0, 0 };
    /* mid = GetPos() = the max corner of slab 0 (:218-223) */
    double mid[3];
    for (int k = 0; k < 3; ++k) mid[k] = sl[0]->max[k];
    int sw = g->slab_order[0], se = g->slab_order[1], nw = g->slab_order[2], ne = g->slab_order[3];
    if (bound_min[0] < curr_min[0] || bound_max[0] > curr_max[0]) {
        g->slab_order[0] = se; g->slab_order[1] = sw; g->slab_order[2] = ne; g->slab_order[3] = nw;
        if (bound_min[0] < curr_min[0]) { mid[0] = g->piece[sw].min[0]; need_update[0] = need_update[2] = 1; }
        else { mid[0] = g->piece[se].max[0]; need_update[1] = need_update[3] = 1; }
    }
    sw = g->slab_order[0]; se = g->slab_order[1]; nw = g->slab_order[2]; ne = g->slab_order[3];
    if (bound_min[2] < curr_min[2] || bound_max[2] > curr_max[2]) {
        g->slab_order[0] = nw; g->slab_order[1] = ne; g->slab_order[2] = sw; g->slab_order[3] = se;
        if (bound_min[2] < curr_min[2]) { mid[2] = g->piece[sw].min[2]; need_update[0] = need_update[1] = 1; }
        else { mid[2] = g->piece[nw].max[2]; need_update[2] = need_update[3] = 1; }
    }
    for (int s = 0; s < 4; ++s) {
        if (!need_update[s]) continue;
        double bmin[3], bmax[3];
        slab_bounds(s, g->ground_width, mid, bmin, bmax);
        build_slab(g, s, bmin, bmax);
    }
}

void trlo_ground3d_view(const trlo_ground_var3d* g, trlo_ground* out)
{
    memset(out, 0, sizeof(*out));
    out->kind = TRLO_GROUND_VAR3D;
    out->num_pieces = 4;
    for (int s = 0; s < 4; ++s) {
        const trlo_terrain_piece* pc = &g->piece[g->slab_order[s]];
        out->data[s] = pc->data;
        out->res[s][0] = pc->res[0];
        out->res[s][1] = pc->res[1];
        for (int k = 0; k < 3; ++k) { out->origin[s][k] = pc->origin[k]; out->scale[s][k] = pc->scale[k]; }
        /* tSlab::ContainsPos :596-600 tests the slab's own double bounds */
        out->bounds[s][0] = pc->min[0];
        out->bounds[s][1] = pc->max[0];
        out->bounds[s][2] = pc->min[2];
        out->bounds[s][3] = pc->max[2];
    }
}
