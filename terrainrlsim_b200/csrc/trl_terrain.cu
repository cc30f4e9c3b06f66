// Procedural terrain on the device: cGroundVar2D / cGroundVar3D with cTerrainGen2D / cTerrainGen3D::
// BuildFlat, generated and scrolled where the sampling kernels read it -- no per-env host work between steps.
//
//   cRand on std::default_random_engine            util/Rand.cpp:6-139
//   cTerrainGen2D::Build* / Add* / Overlay*        sim/TerrainGen2D.cpp:127-653
//   cGroundVar2D::InitSegments / BuildSegment / Update, tSegment::Init   sim/GroundVar2D.cpp:33-96,261-450,457-520
//   cGroundVar3D::InitSlabs / BuildSlab / Update, tSlab::Init            sim/GroundVar3D.cpp:24-141,350-420,511-550
//
// Mapping: ONE THREAD PER TERRAIN. A terrain is a strictly sequential object -- every height depends on the previous one
// and on the position in its own random stream (minstd_rand0 is a serial recurrence) -- and there are thousands of them
// (one per env up to 4096), so the parallel dimension is the terrain; the work is a one-off (init) or rare (a segment is
// rebuilt when a character has advanced 20 m). Generator state (stream position, segment / slab order, extents) stays in a
// small per-terrain record on the device; the kernels write the piece descriptors (DevPiece) the sampling kernels read.
//
// This translation unit is compiled with -fmad=false: heights and extents must be bit-identical to the reference's
// double / float expression sequence (checked against the CPU oracle and the reference-generated golden fixtures), so no
// multiply-add may be contracted.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cfloat>
#include <cmath>

#include "trl_internal.h"

namespace {

// ------------------------------------------------------------------------------------------------
// cRand: std::default_random_engine = minstd_rand0 (x <- 16807 x mod 2^31 - 1), libstdc++'s generate_canonical<double, 53>
// (two draws) and uniform_int_distribution<int> (down-scaling rejection loop), the reference's Linux toolchain
// ------------------------------------------------------------------------------------------------
constexpr unsigned long long kM = 2147483647ULL, kA = 16807ULL;

struct Rand {
    unsigned long long x;
    __device__ unsigned long long next() {
        x = (kA * x) % kM;
        return x;
    }
    __device__ double uniform01() {
        const double R = 2147483646.0;
        double sum = 0.0, tmp = 1.0;
        for (int k = 0; k < 2; ++k) {
            sum += (double)(next() - 1ULL) * tmp;
            tmp *= R;
        }
        double ret = sum / tmp;
        if (ret >= 1.0) ret = nextafter(1.0, 0.0);
        return ret * (1.0 - 0.0) + 0.0;
    }
    __device__ double range(double mn, double mx) {  // cRand::RandDouble(min, max), Rand.cpp:29-40
        if (mn == mx) return mn;
        const double u = uniform01();
        return mn + (u * (mx - mn));
    }
    __device__ unsigned long long below(unsigned long long urange) {  // uniform on [0, urange], generator range [1, m - 1]
        const unsigned long long urngrange = kM - 2ULL;
        unsigned long long ret;
        if (urngrange > urange) {
            const unsigned long long uerange = urange + 1, scaling = urngrange / uerange, past = uerange * scaling;
            do ret = next() - 1ULL; while (ret >= past);
            ret /= scaling;
        } else if (urngrange < urange) {  // up-scaling: one recursion level suffices for 32-bit ranges
            unsigned long long tmp;
            do {
                const unsigned long long uerngrange = urngrange + 1;
                const unsigned long long hi_range = urange / uerngrange;  // < urngrange: the down-scaling branch
                const unsigned long long uer = hi_range + 1, sc = urngrange / uer, past = uer * sc;
                unsigned long long hi;
                do hi = next() - 1ULL; while (hi >= past);
                hi /= sc;
                tmp = uerngrange * hi;
                ret = tmp + (next() - 1ULL);
            } while (ret > urange || ret < tmp);
        } else {
            ret = next() - 1ULL;
        }
        return ret;
    }
    __device__ int rand_int() {  // uniform_int_distribution<int>(INT_MIN + 1, INT_MAX), Rand.cpp:11,57-60
        const long long v = (long long)below(4294967294ULL) + (-2147483647LL);
        return (int)v;
    }
    __device__ int int_range(int mn, int mx) {  // Rand.cpp:62-75
        if (mn == mx) return mn;
        const int v = abs(rand_int());
        return mn + v % (mx - mn);
    }
    __device__ bool flip_coin(double p) { return range(0.0, 1.0) < p; }  // Rand.cpp:136-139
    __device__ int sign() { return flip_coin(0.5) ? -1 : 1; }           // Rand.cpp:131-134
};

// ------------------------------------------------------------------------------------------------
// cTerrainGen2D
// ------------------------------------------------------------------------------------------------
__device__ __constant__ float kSpacing2D = 0.025f;  // TerrainGen2D.cpp:4-10 (SHARP_TERRAIN)

enum { P_GAP_SPACING_MIN, P_GAP_SPACING_MAX, P_GAP_W_MIN, P_GAP_W_MAX, P_GAP_D_MIN, P_GAP_D_MAX,
       P_WALL_SPACING_MIN, P_WALL_SPACING_MAX, P_WALL_W_MIN, P_WALL_W_MAX, P_WALL_H_MIN, P_WALL_H_MAX,
       P_STEP_SPACING_MIN, P_STEP_SPACING_MAX, P_STEP_H0_MIN, P_STEP_H0_MAX, P_STEP_H1_MIN, P_STEP_H1_MAX,
       P_BUMP_H_MIN, P_BUMP_H_MAX,
       P_NGAP_SPACING_MIN, P_NGAP_SPACING_MAX, P_NGAP_DIST_MIN, P_NGAP_DIST_MAX, P_NGAP_W_MIN, P_NGAP_W_MAX,
       P_NGAP_D_MIN, P_NGAP_D_MAX, P_NGAP_COUNT_MIN, P_NGAP_COUNT_MAX,
       P_CLIFF_SPACING_MIN, P_CLIFF_SPACING_MAX, P_CLIFF_H0_MIN, P_CLIFF_H0_MAX, P_CLIFF_H1_MIN, P_CLIFF_H1_MAX,
       P_CLIFF_MINI_COUNT_MAX, P_SLOPE_DELTA_RANGE, P_SLOPE_DELTA_MIN, P_SLOPE_DELTA_MAX, P_MAX };
static_assert(P_MAX == TRL_TERRAIN_NUM_PARAMS, "cTerrainGen2D::eParams");

// the height profile under construction: a fixed-capacity row in global memory (overflow is recorded, never written)
struct Profile {
    float* d;
    int n, cap;
    int overflow;
    __device__ void push(float v) {
        if (n < cap) d[n] = v;
        else overflow = 1;
        ++n;
    }
    __device__ float back() const { return d[(n < cap ? n : cap) - 1]; }
};

struct Gen2D {
    const double* p;  // blended parameters
    Rand* rng;
    Profile* out;

    __device__ static int num_verts(double w) { return (int)ceil(w / kSpacing2D) + 1; }  // :464-468

    __device__ double add_flat(double width) {  // :470-498
        int nv = num_verts(width);
        const int size0 = out->n;
        const bool start_empty = size0 == 0;
        float base_h = 0;
        if (!start_empty) { --nv; base_h = out->back(); }
        for (int i = 0; i < nv; ++i) out->push(base_h);
        int added = out->n - size0;
        if (start_empty) --added;
        return added * kSpacing2D;
    }
    __device__ double add_box(double spacing, double width, double depth) {  // :500-539
        int nv = num_verts(spacing);
        const int size0 = out->n;
        const bool start_empty = size0 == 0;
        float base_h = 0;
        if (!start_empty) { --nv; base_h = out->back(); }
        for (int i = 0; i < nv; ++i) out->push(base_h);
        nv = num_verts(width) - 1;
        const float gap_h = (float)(base_h + depth);
        for (int i = 0; i < nv; ++i) out->push(gap_h);
        out->push(base_h);
        int added = out->n - size0;
        if (start_empty) --added;
        return added * kSpacing2D;
    }
    __device__ double add_step(double width, double height) {  // :541-571
        int nv = num_verts(width);
        const int size0 = out->n;
        const bool start_empty = size0 == 0;
        float base_h = 0;
        if (!start_empty) { --nv; base_h = out->back(); }
        for (int i = 0; i < nv; ++i) out->push(base_h);
        out->push((float)(base_h + height));
        int added = out->n - size0;
        if (start_empty) --added;
        return added * kSpacing2D;
    }
    __device__ void overlay_slopes(double delta_range, double delta_min, double delta_max, double init_slope, int beg, int end) {  // :609-641
        double curr_slope = init_slope, curr_dh = 0.0;
        const double delta_mean = 0.5 * (delta_min + delta_max), delta_diff = 0.5 * (delta_max - delta_min);
        for (int i = beg; i < end; ++i) {
            if (delta_min != delta_max) {
                double delta = rng->range(0.0, delta_range);
                const double sign_rand = rng->range(-1.0, 1.0);
                const double sign_threshold = (curr_slope - delta_mean) / delta_diff;
                delta = (sign_rand < sign_threshold) ? -delta : delta;
                curr_slope += delta;
                curr_dh += curr_slope * kSpacing2D;
            } else {
                curr_dh = 0.0;
            }
            if (i < out->cap) out->d[i] += (float)curr_dh;
        }
    }
    __device__ void overlay_bumps(double min_dh, double max_dh, int beg, int end) {  // :643-653
        for (int i = beg; i < end - 1; ++i) {
            const int sgn = rng->sign();
            const double delta = sgn * rng->range(min_dh, max_dh);
            if (i < out->cap) out->d[i] += (float)delta;
        }
    }
    __device__ void pick_range(double h0_min, double h0_max, double h1_min, double h1_max, double& min_h, double& max_h) {  // :166-186,404-424
        const bool valid_h0 = (h0_min != 0 || h0_max != 0), valid_h1 = (h1_min != 0 || h1_max != 0);
        if (valid_h0 && valid_h1) {
            const bool heads = rng->flip_coin(0.5);
            min_h = heads ? h0_min : h1_min;
            max_h = heads ? h0_max : h1_max;
        } else if (valid_h0) {
            min_h = h0_min; max_h = h0_max;
        } else {
            min_h = h1_min; max_h = h1_max;
        }
    }
    __device__ double gaps(double width) {  // BuildGaps :132-153
        double total = 0.0;
        while (total < width && !out->overflow) {
            const double spacing = rng->range(p[P_GAP_SPACING_MIN], p[P_GAP_SPACING_MAX]);
            const double w = rng->range(p[P_GAP_W_MIN], p[P_GAP_W_MAX]);
            const double d = rng->range(p[P_GAP_D_MIN], p[P_GAP_D_MAX]);
            total += add_box(spacing, w, d);
        }
        return total;
    }
    __device__ double steps(double width) {  // BuildSteps :155-197
        double total = 0.0;
        while (total < width && !out->overflow) {
            double min_h = 0, max_h = 0;
            pick_range(p[P_STEP_H0_MIN], p[P_STEP_H0_MAX], p[P_STEP_H1_MIN], p[P_STEP_H1_MAX], min_h, max_h);
            const double w = rng->range(p[P_STEP_SPACING_MIN], p[P_STEP_SPACING_MAX]);
            const double h = rng->range(min_h, max_h);
            total += add_step(w, h);
        }
        return total;
    }
    __device__ double walls(double width) {  // BuildWalls :199-220
        double total = 0.0;
        while (total < width && !out->overflow) {
            const double spacing = rng->range(p[P_WALL_SPACING_MIN], p[P_WALL_SPACING_MAX]);
            const double w = rng->range(p[P_WALL_W_MIN], p[P_WALL_W_MAX]);
            const double h = rng->range(p[P_WALL_H_MIN], p[P_WALL_H_MAX]);
            total += add_box(spacing, w, h);
        }
        return total;
    }
    __device__ double mixed(double width) {  // BuildMixed :236-264: one feature of a random kind at a time
        double total = 0.0;
        const double one = kSpacing2D;
        while (total < width && !out->overflow) {
            double w = 0.0;
            const int kind = rng->int_range(0, 3);
            if (kind == 0) w = gaps(one);
            else if (kind == 1) w = steps(one);
            else if (kind == 2) w = walls(one);
            total += w;
        }
        return total;
    }
    __device__ double narrow_gaps(double width) {  // BuildNarrowGaps :266-298
        const int count_min = (int)p[P_NGAP_COUNT_MIN] > 1 ? (int)p[P_NGAP_COUNT_MIN] : 1;
        const int count_max = (int)p[P_NGAP_COUNT_MAX] > 1 ? (int)p[P_NGAP_COUNT_MAX] : 1;
        double total = 0.0;
        while (total < width && !out->overflow) {
            double spacing = rng->range(p[P_NGAP_SPACING_MIN], p[P_NGAP_SPACING_MAX]);
            const int count = rng->int_range(count_min, count_max + 1);
            for (int i = 0; i < count; ++i) {
                const double w = rng->range(p[P_NGAP_W_MIN], p[P_NGAP_W_MAX]);
                const double d = rng->range(p[P_NGAP_D_MIN], p[P_NGAP_D_MAX]);
                total += add_box(spacing, w, d);
                spacing = rng->range(p[P_NGAP_DIST_MIN], p[P_NGAP_DIST_MAX]);
            }
        }
        return total;
    }
    __device__ double cliffs(double width) {  // BuildCliffs :392-462
        const int mini_max = (int)p[P_CLIFF_MINI_COUNT_MAX];
        double total = 0.0;
        while (total < width && !out->overflow) {
            double min_h = 0, max_h = 0;
            pick_range(p[P_CLIFF_H0_MIN], p[P_CLIFF_H0_MAX], p[P_CLIFF_H1_MIN], p[P_CLIFF_H1_MAX], min_h, max_h);
            const double w = rng->range(p[P_CLIFF_SPACING_MIN], p[P_CLIFF_SPACING_MAX]);
            const double h = rng->range(min_h, max_h);
            double cw = 0.0, cdh = 0.0;
            const int num_mini = rng->int_range(0, mini_max + 1);
            for (int i = 0; i < num_mini + 1; ++i) {
                const double mini_w = (i == 0) ? w : 0.1;
                double mini_h = rng->range(cdh, h);
                mini_h = (i == num_mini) ? h : mini_h;
                cw += add_step(mini_w, mini_h - cdh);
                cdh = mini_h;
            }
            total += cw;
        }
        return total;
    }
    // cTerrainGen2D::GetTerrainFunc :86-123 and the Build* wrappers :127-131,222-234,300-390; types = cGround::eType
    __device__ double base(int type, double width) {
        switch (type) {
        case 2: return gaps(width);
        case 3: return steps(width);
        case 4: return walls(width);
        case 6: return mixed(width);
        case 7: return narrow_gaps(width);
        case 14: return cliffs(width);
        default: return add_flat(width);
        }
    }
    __device__ double build(int type, double width) {
        if (type == 5) {  // var2d_bumps
            const int beg = out->n;
            const double total = add_flat(width);
            overlay_bumps(p[P_BUMP_H_MIN], p[P_BUMP_H_MAX], beg, out->n);
            return total;
        }
        if (type >= 8 && type <= 13) {  // var2d_slopes*: a base terrain with a random-walk slope laid over it
            const int base_type = (type == 8) ? 1 : (type == 9) ? 2 : (type == 10) ? 4 : (type == 11) ? 3 : (type == 12) ? 6 : 7;
            const int beg = out->n;
            const double total = base(base_type, width);
            overlay_slopes(fabs(p[P_SLOPE_DELTA_RANGE]), p[P_SLOPE_DELTA_MIN], p[P_SLOPE_DELTA_MAX], 0.0, beg, out->n);
            return total;
        }
        return base(type, width);
    }
};

// ------------------------------------------------------------------------------------------------
// Bullet-side float round trips of a heightfield body (cWorld::SetPos / GetPos sim/World.cpp:333-362, setLocalScaling,
// btHeightfieldTerrainShape::getAabb): origin, spacing and the rigid body's x / z extent as the reference reads them back
// ------------------------------------------------------------------------------------------------
__device__ void place_piece(DevPiece& pc, double aabb[4], const double origin[3], double x_scale, double z_scale,
                            double world_scale, int res_x, int res_z) {
    const float sc = (float)world_scale;
    for (int k = 0; k < 3; ++k) {
        const float o = sc * (float)origin[k];
        pc.origin[k] = (double)o / world_scale;
    }
    const float ls[3] = {(float)x_scale, 1.0f, (float)z_scale};
    for (int k = 0; k < 3; ++k) pc.scale[k] = (double)ls[k] / world_scale;
    const float hx = ((float)(res_x - 1) - 0.0f) * ls[0] * 0.5f, hz = ((float)(res_z - 1) - 0.0f) * ls[2] * 0.5f;
    const float cx = sc * (float)origin[0], cz = sc * (float)origin[2];
    aabb[0] = (double)(cx - hx) / world_scale;
    aabb[1] = (double)(cx + hx) / world_scale;
    aabb[2] = (double)(cz - hz) / world_scale;
    aabb[3] = (double)(cz + hz) / world_scale;
    pc.res_x = res_x;
    pc.res_z = res_z;
    pc.rscale[0] = 1.0 / pc.scale[0];  // what trl_ground_set_heightfields derives on the host (IEEE division either way)
    pc.rscale[1] = 1.0 / pc.scale[2];
    pc.half[0] = (double)(res_x - 1) * 0.5;
    pc.half[1] = (double)(res_z - 1) * 0.5;
    pc.cmax3[0] = res_x - 1.0 - 0.0001;
    pc.cmax3[1] = res_z - 1.0 - 0.0001;
    pc.pad_[0] = pc.pad_[1] = pc.pad_[2] = 0;
}

// ------------------------------------------------------------------------------------------------
// cGroundVar2D: two segments [min | max]
// ------------------------------------------------------------------------------------------------
enum { ALIGN_MIN = 0, ALIGN_MAX = 1 };

struct Ground2D {
    const TerrainGenCfg& cfg;
    TerrainState& st;
    float* slots;      // this terrain's two profile rows, cfg.slot_floats apart
    long long slot0;   // offset of slot 0 in the device heightfield array
    DevPiece* view;    // [2] descriptors in the reference's scan order (GetSegment(s))
    Rand rng;

    __device__ int seg_id(int s) const { return st.flip ? (s == 0 ? 1 : 0) : s; }  // :316-323
    __device__ int seg_align(int s) const { return st.flip ? (s == 0 ? ALIGN_MIN : ALIGN_MAX) : (s == 0 ? ALIGN_MAX : ALIGN_MIN); }  // :349-356
    __device__ double min_x(int id) const { return st.count[id] == 0 ? CUDART_INF : st.aabb[id][0]; }
    __device__ double max_x(int id) const { return st.count[id] == 0 ? -CUDART_INF : st.aabb[id][1]; }

    // cGroundVar2D::BuildSegment :358-388 (+ AddPadding :390-397) and tSegment::Init :457-520
    __device__ void build_segment(int id, double bound_min, double bound_max, int align, const double fix_point[2]) {
        Profile prof{slots + (size_t)id * cfg.slot_floats, 0, (int)cfg.slot_floats, 0};
        Gen2D gen{cfg.params, &rng, &prof};
        const bool contains_origin = (bound_min <= 0) && (bound_max >= 0);
        if (contains_origin) gen.add_flat(fmin(bound_max - bound_min, 1 - bound_min));
        gen.build(cfg.type, bound_max - bound_min);
        if (prof.overflow) { st.overflow = 1; prof.n = prof.cap; }
        const int nv = prof.n;
        float end_h = 0;
        if (nv > 0) end_h = (align == ALIGN_MIN) ? prof.d[0] : prof.d[nv - 1];
        const float h_offset = (float)(fix_point[1] - end_h);
        const double ws = cfg.world_scale;
        const double new_min = (align == ALIGN_MIN) ? bound_min : (bound_max - (nv - 1) * (double)kSpacing2D);
        double lo[3] = {new_min, CUDART_INF, 0}, hi[3] = {new_min + (nv - 1) * (double)kSpacing2D, -CUDART_INF, 0};
        for (int i = 0; i < nv; ++i) {
            const float v = prof.d[i] + h_offset;
            const double h = v;
            if (h < lo[1]) lo[1] = h;
            if (h > hi[1]) hi[1] = h;
            prof.d[i] = v * (float)ws;
        }
        st.count[id] = nv;
        double origin[3];
        for (int k = 0; k < 3; ++k) origin[k] = 0.5 * (lo[k] + hi[k]);
        DevPiece& pc = st.piece[id];
        place_piece(pc, st.aabb[id], origin, (double)kSpacing2D * ws, 1.0 * ws, ws, nv, 3);
        pc.data_offset = slot0 + (long long)id * cfg.slot_floats;
        pc.pitch = nv;
        // ContainsPos :697-700: GetMinX / GetMaxX = the rigid body's AABB (:540-564); z is not tested
        pc.bounds[0] = st.aabb[id][0];
        pc.bounds[1] = st.aabb[id][1];
        pc.bounds[2] = -CUDART_INF;
        pc.bounds[3] = CUDART_INF;
    }
    __device__ void init_segments(double bound_min_x, double bound_max_x) {  // :290-314
        st.count[0] = st.count[1] = 0;
        st.flip = 0;
        const double mid = 0.5 * (bound_max_x + bound_min_x), w = cfg.ground_width;
        for (int i = 0; i < 2; ++i) {
            const double fix_point[2] = {0, 0};
            const int align = seg_align(i);
            double bmin = (align == ALIGN_MAX) ? -w : 0, bmax = (align == ALIGN_MAX) ? 0 : w;
            bmin += mid;
            bmax += mid;
            build_segment(seg_id(i), bmin, bmax, align, fix_point);
        }
    }
    __device__ bool update(double bound_min_x, double bound_max_x) {  // cGroundVar2D::Update :48-96
        const int min_id = seg_id(0), max_id = seg_id(1);
        const double mn = min_x(min_id), mx = max_x(max_id);
        if (bound_max_x < mx && bound_min_x > mn) return false;
        if (bound_max_x <= mn || bound_min_x >= mx) {
            init_segments(bound_min_x, bound_max_x);
            return true;
        }
        int id, align;
        double seg_min, seg_max, fix_point[2] = {0, 0};
        const float* rows = slots;
        if (bound_max_x >= mx) {
            id = seg_id(0);
            seg_min = mx;
            seg_max = seg_min + cfg.ground_width;
            fix_point[0] = mx;
            fix_point[1] = rows[(size_t)max_id * cfg.slot_floats + st.count[max_id] - 1] / cfg.world_scale;  // GetEndHeight :646-650
            align = ALIGN_MIN;
        } else {
            id = seg_id(1);
            seg_max = mn;
            seg_min = seg_max - cfg.ground_width;
            fix_point[0] = mn;
            fix_point[1] = rows[(size_t)min_id * cfg.slot_floats] / cfg.world_scale;  // GetStartHeight :640-644
            align = ALIGN_MAX;
        }
        build_segment(id, seg_min, seg_max, align, fix_point);
        st.flip = (seg_id(0) == 0);
        return true;
    }
    __device__ void publish() {
        for (int s = 0; s < 2; ++s) view[s] = st.piece[seg_id(s)];
    }
};

__global__ void __launch_bounds__(64) k_terrain2d(TerrainGenCfg cfg, int F, const unsigned long long* __restrict__ seeds,
                                                  const double* __restrict__ bmin, const double* __restrict__ bmax,
                                                  TerrainState* __restrict__ states, float* __restrict__ data,
                                                  DevPiece* __restrict__ pieces, int* __restrict__ changed, int init) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    TerrainState& st = states[f];
    const long long slot0 = (long long)f * 2 * cfg.slot_floats;
    Ground2D g{cfg, st, data + slot0, slot0, pieces + (size_t)f * 2, Rand{0}};
    bool touched;
    if (init) {
        st.overflow = 0;
        const unsigned long long s = seeds[f] % kM;  // linear_congruential_engine::seed: 0 mod m becomes 1
        g.rng.x = (s == 0) ? 1 : s;
        g.init_segments(bmin[3 * f], bmax[3 * f]);
        touched = true;
    } else {
        g.rng.x = st.rand_x;
        touched = g.update(bmin[3 * f], bmax[3 * f]);
    }
    st.rand_x = g.rng.x;
    if (touched) g.publish();
    if (changed) changed[f] = st.overflow ? 2 : (touched ? 1 : 0);
}

// ------------------------------------------------------------------------------------------------
// cGroundVar3D with cTerrainGen3D::BuildFlat: four slabs [SW SE | NW NE]; one thread per terrain decides, then the block's
// threads fill the rebuilt slabs (a slab is 201 x 201 floats)
// ------------------------------------------------------------------------------------------------
__device__ int res_3d(double size, double spacing) { return (int)ceil(size / spacing - 0.000001) + 1; }  // TerrainGen3D.cpp:70-78

__device__ void slab_bounds(int s, double w, const double mid[3], double bmin[3], double bmax[3]) {  // GroundVar3D.cpp:358-369
    bmin[0] = (s % 2 == 0) ? -w : 0;
    bmax[0] = (s % 2 == 0) ? 0 : w;
    bmin[1] = bmax[1] = 0;
    bmin[2] = (s < 2) ? -w : 0;
    bmax[2] = (s < 2) ? 0 : w;
    for (int k = 0; k < 3; ++k) { bmin[k] += mid[k]; bmax[k] += mid[k]; }
}

// descriptor and extents of slab position s (BuildSlab :389-414, tSlab::Init :511-550); the heights (BuildFlat: 0, times the
// world scale) are written by the whole block afterwards
__device__ void describe_slab(const TerrainGenCfg& cfg, TerrainState& st, long long slot0, int s, const double bmin[3], const double bmax[3]) {
    const int id = st.order[s];
    const double size_x = bmax[0] - bmin[0], size_z = bmax[2] - bmin[2];
    const int res_x = res_3d(size_x, cfg.spacing_x), res_z = res_3d(size_z, cfg.spacing_z);
    const int pitch = (res_x + 3) & ~3;
    if ((long long)pitch * res_z > cfg.slot_floats) { st.overflow = 1; return; }
    double mn[3] = {bmin[0], bmin[1], bmin[2]}, mx[3] = {bmax[0], bmax[1], bmax[2]};
    const double h = (float)0.0;  // every vertex of a flat slab
    if (h < mn[1]) mn[1] = h;
    if (h > mx[1]) mx[1] = h;
    for (int k = 0; k < 3; ++k) { st.mn[id][k] = mn[k]; st.mx[id][k] = mx[k]; }
    double origin[3];
    for (int k = 0; k < 3; ++k) origin[k] = 0.5 * (mx[k] + mn[k]);
    DevPiece& pc = st.piece[id];
    const double ws = cfg.world_scale;
    place_piece(pc, st.aabb[id], origin, cfg.spacing_x * ws, cfg.spacing_z * ws, ws, res_x, res_z);
    pc.data_offset = slot0 + (long long)id * cfg.slot_floats;
    pc.pitch = pitch;
    pc.bounds[0] = mn[0];  // tSlab::ContainsPos :596-600 tests the slab's own double bounds
    pc.bounds[1] = mx[0];
    pc.bounds[2] = mn[2];
    pc.bounds[3] = mx[2];
    st.count[id] = res_x * res_z;
}

__global__ void __launch_bounds__(128) k_terrain3d(TerrainGenCfg cfg, int F, const double* __restrict__ bound_min,
                                                   const double* __restrict__ bound_max, TerrainState* __restrict__ states,
                                                   float* __restrict__ data, DevPiece* __restrict__ pieces,
                                                   int* __restrict__ changed, int init) {
    const int f = blockIdx.x;
    if (f >= F) return;
    __shared__ int s_rebuild;  // bit s: slab position s is rebuilt
    TerrainState& st = states[f];
    const long long slot0 = (long long)f * 4 * cfg.slot_floats;
    if (threadIdx.x == 0) {
        const double* bmin_w = bound_min + 3 * (size_t)f;
        const double* bmax_w = bound_max + 3 * (size_t)f;
        int rebuild = 0;
        double mid[3];
        bool all = init != 0;
        if (init) {
            st.overflow = 0;
            for (int s = 0; s < 4; ++s) st.order[s] = s;
        } else {  // cGroundVar3D::Update :38-141
            const int o0 = st.order[0], o1 = st.order[1], o2 = st.order[2], o3 = st.order[3];
            const double cmin[3] = {fmax(st.mn[o0][0], st.mn[o2][0]), 0, fmax(st.mn[o0][2], st.mn[o1][2])};  // GetBounds :422-438
            const double cmax[3] = {fmin(st.mx[o1][0], st.mx[o3][0]), 0, fmin(st.mx[o2][2], st.mx[o3][2])};
            if (bmin_w[0] > cmin[0] && bmin_w[2] > cmin[2] && bmax_w[0] < cmax[0] && bmax_w[2] < cmax[2]) {
                // inside: nothing to do
            } else if ((bmin_w[0] > cmax[0] && bmin_w[2] > cmax[2]) || (bmax_w[0] < cmin[0] && bmax_w[2] < cmin[2])) {
                all = true;
            } else {
                for (int k = 0; k < 3; ++k) mid[k] = st.mx[o0][k];  // GetPos() = the max corner of slab 0 (:218-223)
                int sw = st.order[0], se = st.order[1], nw = st.order[2], ne = st.order[3];
                if (bmin_w[0] < cmin[0] || bmax_w[0] > cmax[0]) {
                    st.order[0] = se; st.order[1] = sw; st.order[2] = ne; st.order[3] = nw;
                    if (bmin_w[0] < cmin[0]) { mid[0] = st.mn[sw][0]; rebuild |= 1 | 4; }
                    else { mid[0] = st.mx[se][0]; rebuild |= 2 | 8; }
                }
                sw = st.order[0]; se = st.order[1]; nw = st.order[2]; ne = st.order[3];
                if (bmin_w[2] < cmin[2] || bmax_w[2] > cmax[2]) {
                    st.order[0] = nw; st.order[1] = ne; st.order[2] = sw; st.order[3] = se;
                    if (bmin_w[2] < cmin[2]) { mid[2] = st.mn[sw][2]; rebuild |= 1 | 2; }
                    else { mid[2] = st.mx[nw][2]; rebuild |= 4 | 8; }
                }
            }
        }
        if (all) {  // InitSlabs :350-372
            for (int k = 0; k < 3; ++k) mid[k] = 0.5 * (bmax_w[k] + bmin_w[k]);
            rebuild = 15;
        }
        for (int s = 0; s < 4; ++s)
            if ((rebuild >> s) & 1) {
                double bmin[3], bmax[3];
                slab_bounds(s, cfg.ground_width, mid, bmin, bmax);
                describe_slab(cfg, st, slot0, s, bmin, bmax);
            }
        if (rebuild || init) {
            for (int s = 0; s < 4; ++s) pieces[(size_t)f * 4 + s] = st.piece[st.order[s]];
        }
        if (changed) changed[f] = st.overflow ? 2 : (rebuild ? 1 : 0);
        s_rebuild = st.overflow ? 0 : rebuild;
    }
    __syncthreads();
    const int rebuild = s_rebuild;
    for (int s = 0; s < 4; ++s) {
        if (!((rebuild >> s) & 1)) continue;
        const DevPiece& pc = st.piece[st.order[s]];
        float* dst = data + pc.data_offset;
        const long long count = (long long)pc.pitch * pc.res_z;
        const float v = (float)0.0 * (float)cfg.world_scale;  // AddFlat (TerrainGen3D.cpp:182-203), scaled like tSlab::Init
        for (long long i = threadIdx.x; i < count; i += blockDim.x) dst[i] = v;
    }
}

}  // namespace

int trl_launch_terrain(const TerrainGenCfg& cfg, int F, const unsigned long long* d_seeds, const double* d_bound_min,
                       const double* d_bound_max, TerrainState* d_states, float* d_data, DevPiece* d_pieces, int* d_changed,
                       int init, void* stream) {
    if (cfg.type == TRL_TERRAIN_VAR3D_FLAT) {
        k_terrain3d<<<F, 128, 0, (cudaStream_t)stream>>>(cfg, F, d_bound_min, d_bound_max, d_states, d_data, d_pieces, d_changed, init);
    } else {
        k_terrain2d<<<(F + 63) / 64, 64, 0, (cudaStream_t)stream>>>(cfg, F, d_seeds, d_bound_min, d_bound_max, d_states, d_data,
                                                                    d_pieces, d_changed, init);
    }
    return (int)cudaGetLastError();
}
