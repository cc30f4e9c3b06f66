// Host-side model loader: reference character / motion JSON and arg files -> flat DevModel.
// Mirrors the *data formats* of the reference loaders (cited per function); the parser is our own.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>

#include "trl_internal.h"

// ------------------------------------------------------------------------------------------------
// error string (thread local, like errno)
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
void trl_set_error(const std::string& msg) { g_last_error = msg; }
extern "C" const char* trl_last_error(void) { return g_last_error.c_str(); }

// ------------------------------------------------------------------------------------------------
// minimal JSON reader (objects, arrays, strings, numbers, true/false/null, // and /* */ comments,
// which the reference's jsoncpp reader also tolerates)
// ------------------------------------------------------------------------------------------------
namespace {

struct JVal {
    enum Kind { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
    double num = 0;
    bool b = false;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;

    const JVal* get(const std::string& key) const {
        if (kind != OBJ) return nullptr;
        for (const auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_null() const { return kind == NUL; }
    bool is_numeric() const { return kind == NUM || kind == BOOL; }  // Json::Value::isNumeric counts bools
    double as_double() const { return kind == NUM ? num : (kind == BOOL ? (b ? 1.0 : 0.0) : 0.0); }
};

class JParser {
public:
    explicit JParser(const std::string& s) : s_(s) {}
    bool parse(JVal& out, std::string& err) {
        try {
            skip();
            out = value();
            skip();
            if (p_ != s_.size()) throw std::string("trailing characters");
            return true;
        } catch (const std::string& e) {
            std::ostringstream os;
            os << e << " at byte " << p_;
            err = os.str();
            return false;
        }
    }

private:
    const std::string& s_;
    size_t p_ = 0;

    void skip() {
        for (;;) {
            while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\t' || s_[p_] == '\n' || s_[p_] == '\r')) ++p_;
            if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '/') {
                while (p_ < s_.size() && s_[p_] != '\n') ++p_;
            } else if (p_ + 1 < s_.size() && s_[p_] == '/' && s_[p_ + 1] == '*') {
                p_ += 2;
                while (p_ + 1 < s_.size() && !(s_[p_] == '*' && s_[p_ + 1] == '/')) ++p_;
                p_ += 2;
            } else {
                return;
            }
        }
    }
    char peek() const { return p_ < s_.size() ? s_[p_] : '\0'; }
    void expect(char c) {
        if (peek() != c) throw std::string("expected '") + c + "'";
        ++p_;
    }
    JVal value() {
        skip();
        char c = peek();
        if (c == '{') return object();
        if (c == '[') return array();
        if (c == '"') {
            JVal v;
            v.kind = JVal::STR;
            v.str = string();
            return v;
        }
        if (s_.compare(p_, 4, "true") == 0) { p_ += 4; JVal v; v.kind = JVal::BOOL; v.b = true; return v; }
        if (s_.compare(p_, 5, "false") == 0) { p_ += 5; JVal v; v.kind = JVal::BOOL; v.b = false; return v; }
        if (s_.compare(p_, 4, "null") == 0) { p_ += 4; return JVal(); }
        return number();
    }
    JVal number() {
        const char* start = s_.c_str() + p_;
        char* end = nullptr;
        JVal v;
        v.kind = JVal::NUM;
        if (s_.compare(p_, 8, "Infinity") == 0) { p_ += 8; v.num = std::numeric_limits<double>::infinity(); return v; }
        if (s_.compare(p_, 9, "-Infinity") == 0) { p_ += 9; v.num = -std::numeric_limits<double>::infinity(); return v; }
        v.num = std::strtod(start, &end);
        if (end == start) throw std::string("bad number");
        p_ += static_cast<size_t>(end - start);
        return v;
    }
    std::string string() {
        expect('"');
        std::string out;
        while (p_ < s_.size() && s_[p_] != '"') {
            if (s_[p_] == '\\' && p_ + 1 < s_.size()) {
                char e = s_[p_ + 1];
                switch (e) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                default: out += e; break;
                }
                p_ += 2;
            } else {
                out += s_[p_++];
            }
        }
        expect('"');
        return out;
    }
    JVal array() {
        JVal v;
        v.kind = JVal::ARR;
        expect('[');
        skip();
        if (peek() == ']') { ++p_; return v; }
        for (;;) {
            v.arr.push_back(value());
            skip();
            if (peek() == ',') { ++p_; skip(); if (peek() == ']') { ++p_; return v; } continue; }
            expect(']');
            return v;
        }
    }
    JVal object() {
        JVal v;
        v.kind = JVal::OBJ;
        expect('{');
        skip();
        if (peek() == '}') { ++p_; return v; }
        for (;;) {
            skip();
            std::string k = string();
            skip();
            expect(':');
            JVal val = value();
            v.obj.emplace_back(std::move(k), std::move(val));
            skip();
            if (peek() == ',') { ++p_; skip(); if (peek() == '}') { ++p_; return v; } continue; }
            expect('}');
            return v;
        }
    }
};

bool read_file(const char* path, std::string& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    std::ostringstream ss;
    ss << f.rdbuf();
    out = ss.str();
    return true;
}

const char* kJointTypeNames[] = {"revolute", "planar", "prismatic", "fixed", "spherical", "none"};
// key order = column order, anim/KinTree.cpp:25-46 (the last column's JSON key is "Offset")
const char* kJointKeys[TRL_JOINT_COLS] = {"Type", "Parent", "AttachX", "AttachY", "AttachZ", "AttachThetaX",
                                          "AttachThetaY", "AttachThetaZ", "LimLow0", "LimLow1", "LimLow2", "LimHigh0",
                                          "LimHigh1", "LimHigh2", "TorqueLim", "ForceLim", "IsEndEffector",
                                          "DiffWeight", "Offset"};
// anim/KinTree.cpp:49-68
const char* kBodyKeys[TRL_BODY_COLS] = {"Shape", "Mass", "ColGroup", "EnableFallContact", "AttachX", "AttachY",
                                        "AttachZ", "AttachThetaX", "AttachThetaY", "AttachThetaZ", "Param0", "Param1",
                                        "Param2", "ColorR", "ColorG", "ColorB", "ColorA"};
// sim/PDController.cpp:8-27
const char* kPDKeys[TRL_PD_COLS] = {"JointID", "Kp", "Kd", "TargetTheta0", "TargetTheta1", "TargetTheta2",
                                    "TargetTheta3", "TargetTheta4", "TargetTheta5", "TargetTheta6", "TargetVel0",
                                    "TargetVel1", "TargetVel2", "TargetVel3", "TargetVel4", "TargetVel5", "TargetVel6",
                                    "UseWorldCoord"};

int joint_param_size(int type) {  // cKinTree::GetJointParamSize, anim/KinTree.cpp:797-823
    switch (type) {
    case JT_REVOLUTE: return 1;
    case JT_PRISMATIC: return 1;
    case JT_PLANAR: return 3;
    case JT_FIXED: return 0;
    case JT_SPHERICAL: return 4;
    default: return -1;
    }
}

// cMathUtil::RotateMat(euler) = Rz*Ry*Rx, util/MathUtil.cpp:128-155
void euler_to_mat(const double e[3], double R[9]) {
    double xs = std::sin(e[0]), xc = std::cos(e[0]);
    double ys = std::sin(e[1]), yc = std::cos(e[1]);
    double zs = std::sin(e[2]), zc = std::cos(e[2]);
    R[0] = yc * zc; R[3] = yc * zs; R[6] = -ys;
    R[1] = xs * ys * zc - xc * zs; R[4] = xs * ys * zs + xc * zc; R[7] = xs * yc;
    R[2] = xc * ys * zc + xs * zs; R[5] = xc * ys * zs - xs * zc; R[8] = xc * yc;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// flatten to DevModel
// ------------------------------------------------------------------------------------------------
int trl_build_dev_model(trl_model* m) {
    DevModel& d = m->dev;
    std::memset(&d, 0, sizeof(d));
    const int N = m->N;
    if (N < 1 || N > TRL_MAX_JOINTS) {
        trl_set_error("model has too many joints (max 32)");
        return TRL_ERR_UNSUPPORTED;
    }
    d.N = N;
    int off = 0;
    for (int j = 0; j < N; ++j) {
        const double* jr = &m->joint_mat[(size_t)j * TRL_JOINT_COLS];
        const double* br = &m->body_defs[(size_t)j * TRL_BODY_COLS];
        const double* pr = &m->pd_params[(size_t)j * TRL_PD_COLS];
        int type = (int)jr[0];
        int parent = (int)jr[1];
        if (j == 0 && parent != -1) { trl_set_error("joint 0 must be the root (Parent = -1)"); return TRL_ERR_UNSUPPORTED; }
        if (j > 0 && (parent < 0 || parent >= j)) {  // anim/KinTree.cpp:479-493
            trl_set_error("Parent id must be < child id");
            return TRL_ERR_UNSUPPORTED;
        }
        int size = (parent == -1) ? 7 : joint_param_size(type);  // root is always 7 (KinTree.cpp:789-795)
        if (size < 0) { trl_set_error("unsupported joint type"); return TRL_ERR_UNSUPPORTED; }
        d.ix.parent[j] = parent;
        d.ix.type[j] = type;
        d.ix.offset[j] = off;
        d.ix.size[j] = size;
        off += size;
        int shape = (int)br[0];
        d.ix.valid_body[j] = shape < 2;  // cKinTree::IsValidBody, KinTree.cpp:390-398
        d.ix.pd_valid[j] = (parent != -1) && d.ix.valid_body[j];
        euler_to_mat(jr + 5, d.attR[j]);
        d.attT[j][0] = jr[2]; d.attT[j][1] = jr[3]; d.attT[j][2] = jr[4];
        if (parent == -1) d.attT[j][0] = d.attT[j][1] = d.attT[j][2] = 0;  // PostProcessJointMat :1038-1040
        d.ix.level[j] = (parent == -1) ? 0 : d.ix.level[parent] + 1;
        d.torque_lim[j] = jr[14];
        d.force_lim[j] = jr[15];
        d.mass[j] = 0;
        if (d.ix.valid_body[j]) {
            double mass = br[1];
            d.mass[j] = mass;
            d.com[j][0] = br[4]; d.com[j][1] = br[5]; d.com[j][2] = br[6];
            if (shape == 0) {  // box, RBDUtil.cpp:623-644
                double sx = br[10], sy = br[11], sz = br[12];
                d.idiag[j][0] = mass / 12.0 * (sy * sy + sz * sz);
                d.idiag[j][1] = mass / 12.0 * (sx * sx + sz * sz);
                d.idiag[j][2] = mass / 12.0 * (sx * sx + sy * sy);
            } else {  // capsule, RBDUtil.cpp:646-673 (with the cm*cm term)
                double r = br[10], h = br[11];
                double c_vol = M_PI * r * r * h;
                double hs_vol = M_PI * 2.0 / 3.0 * r * r * r;
                double density = mass / (c_vol + 2 * hs_vol);
                double cm = c_vol * density;
                double hsm = hs_vol * density;
                double x = cm * (0.25 * r * r + (1.0 / 12.0) * cm * h * h) +
                           2 * hsm * (0.4 * r * r + 0.375 * r * h + 0.25 * h * h);
                d.idiag[j][0] = x;
                d.idiag[j][1] = (0.5 * cm + 0.8 * hsm) * r * r;
                d.idiag[j][2] = x;
            }
        }
        d.ix.end_eff[j] = jr[16] != 0;
        d.kp[j] = d.ix.pd_valid[j] ? pr[1] : 0.0;  // cImpPDController::InitGains, ImpPDController.cpp:100-121
        d.kd[j] = d.ix.pd_valid[j] ? pr[2] : 0.0;
        if (type == JT_SPHERICAL && parent != -1) d.has_spherical = 1;
    }
    d.n = off;
    m->n = off;
    {
        double wsum = 0;
        d.total_mass = 0;
        for (int j = 0; j < N; ++j) {
            wsum += std::fabs(m->joint_mat[(size_t)j * TRL_JOINT_COLS + 17]);
            if (d.ix.valid_body[j]) d.total_mass += d.mass[j];
        }
        for (int j = 0; j < N; ++j) d.joint_w[j] = m->joint_mat[(size_t)j * TRL_JOINT_COLS + 17] / wsum;
    }
    if (off > TRL_MAX_DOF) { trl_set_error("model has too many dofs (max 64)"); return TRL_ERR_UNSUPPORTED; }

    // live dof columns
    int nl = 0;
    for (int i = 0; i < TRL_MAX_DOF; ++i) d.ix.dof_col[i] = -1;
    for (int j = 0; j < N; ++j) {
        d.ix.first_col[j] = nl;
        int o = d.ix.offset[j];
        auto add = [&](int kind, int axis, int dof) {
            d.ix.col_joint[nl] = j; d.ix.col_kind[nl] = kind; d.ix.col_axis[nl] = axis; d.ix.col_dof[nl] = dof;
            d.ix.dof_col[dof] = nl;
            ++nl;
        };
        if (d.ix.parent[j] == -1) {
            for (int k = 0; k < 3; ++k) add(CK_ROOT_LIN, k, o + k);
            for (int k = 0; k < 3; ++k) add(CK_ANG, k, o + 3 + k);
        } else {
            switch (d.ix.type[j]) {
            case JT_REVOLUTE: add(CK_ANG, 2, o); break;
            case JT_PRISMATIC: add(CK_LIN, 0, o); break;
            case JT_PLANAR: add(CK_LIN, 0, o); add(CK_LIN, 1, o + 1); add(CK_ANG, 2, o + 2); break;
            case JT_SPHERICAL: for (int k = 0; k < 3; ++k) add(CK_ANG, k, o + k); break;
            default: break;
            }
        }
        d.ix.ncol[j] = nl - d.ix.first_col[j];
        for (int i = 0; i < d.ix.size[j]; ++i) d.ix.dof_joint[o + i] = j;
    }
    d.nl = nl;
    for (int j = 0; j < N; ++j) {
        double* c = d.pd_cst + 21 * j;
        for (int k = 0; k < 9; ++k) c[k] = d.attR[j][k];
        for (int k = 0; k < 3; ++k) { c[9 + k] = d.attT[j][k]; c[13 + k] = d.com[j][k]; c[16 + k] = d.idiag[j][k]; }
        c[12] = d.mass[j];
        c[19] = d.kp[j];
        c[20] = d.kd[j];
    }
    for (int c = 0; c < nl; ++c) {
        int j = d.ix.col_joint[c];
        if (c > d.ix.first_col[j]) {
            d.ix.col_parent[c] = c - 1;
        } else {
            int p = d.ix.parent[j];
            while (p != -1 && d.ix.ncol[p] == 0) p = d.ix.parent[p];
            d.ix.col_parent[c] = (p == -1) ? -1 : d.ix.first_col[p] + d.ix.ncol[p] - 1;
        }
    }
    {   // branch-sparse H layout of k_pd_block
        PdBlkIdx& b = d.bx;
        std::memset(&b, 0, sizeof(b));
        int total = 0, maxd = 0;
        for (int c = 0; c < nl; ++c) {
            const int pc = d.ix.col_parent[c];
            b.cdepth[c] = (pc == -1) ? 0 : (signed char)(b.cdepth[pc] + 1);
            b.rowbase[c] = (short)total;
            total += b.cdepth[c] + 1;
            maxd = std::max(maxd, (int)b.cdepth[c]);
        }
        b.hcount = total;
        b.max_cdepth = maxd;
        int pos2 = 0;
        for (int dep = 0; dep <= maxd; ++dep) {
            b.cd_start[dep] = (signed char)pos2;
            for (int c = 0; c < nl; ++c)
                if (b.cdepth[c] == dep) b.cd_order[pos2++] = (signed char)c;
        }
        b.cd_start[maxd + 1] = (signed char)pos2;
        int slots = 0;
        b.all_valid = 1;
        for (int j = 0; j < N; ++j) {
            b.use_world[j] = (d.ix.pd_valid[j] && m->pd_params[(size_t)j * TRL_PD_COLS + 17] != 0) ? 1 : 0;
            b.aba_off[j] = (short)slots;
            if (d.ix.parent[j] != -1) slots += 1 + 14 * d.ix.ncol[j] + (d.ix.size[j] - d.ix.ncol[j]);  // + dead slots
            if (!d.ix.valid_body[j]) b.all_valid = 0;
        }
        b.aba_slots = slots;
    }
    // levels
    int max_level = 0;
    for (int j = 0; j < N; ++j) max_level = std::max(max_level, (int)d.ix.level[j]);
    d.num_levels = max_level + 1;
    int pos = 0;
    for (int l = 0; l <= max_level; ++l) {
        d.ix.level_start[l] = pos;
        for (int j = 0; j < N; ++j)
            if (d.ix.level[j] == l) d.ix.level_joint[pos++] = j;
    }
    d.ix.level_start[max_level + 1] = pos;
    {   // depth-first event list (enter / leave) for the thread-per-env tree walk of k_pd_tree
        int ne = 0;
        std::vector<int> stack;
        std::vector<int> next_child(N, 0);
        for (int r = 0; r < N; ++r) {
            if (d.ix.parent[r] != -1) continue;
            stack.push_back(r);
            d.ix.dfs_event[ne++] = (signed char)r;
            while (!stack.empty()) {
                const int j = stack.back();
                int c = next_child[j];
                while (c < N && d.ix.parent[c] != j) ++c;
                if (c < N) {
                    next_child[j] = c + 1;
                    stack.push_back(c);
                    d.ix.dfs_event[ne++] = (signed char)c;
                } else {
                    next_child[j] = N;
                    stack.pop_back();
                    d.ix.dfs_event[ne++] = (signed char)(-1 - j);
                }
            }
        }
    }
    {   // pre-order list, branches (subtrees of the root's children) and their workspace layout for k_pd_aba
        PdBlkIdx& b = d.bx;
        int np = 0;
        for (int k = 0; k < 2 * N; ++k)
            if (d.ix.dfs_event[k] >= 0) b.preorder[np++] = d.ix.dfs_event[k];
        for (int j = 0; j < N; ++j) { b.nchild[j] = 0; b.xoff[j] = -1; }
        for (int j = 1; j < N; ++j) b.nchild[d.ix.parent[j]]++;
        int nb = 0, area = 0;
        for (int k = 1; k < 2 * N - 1;) {  // events between "enter root" and "leave root"
            const int r = d.ix.dfs_event[k];  // enters a child of the root
            int k1 = k;
            while (d.ix.dfs_event[k1] != -1 - r) ++k1;
            b.b_ev0[nb] = (signed char)k;
            b.b_ev1[nb] = (signed char)(k1 + 1);
            int depth = 0, extra = 0, first = -1, cnt = 0;
            for (int q = k; q <= k1; ++q) {
                const int jj = d.ix.dfs_event[q];
                if (jj < 0) continue;
                depth = std::max(depth, (int)d.ix.level[jj]);
                ++cnt;
            }
            for (int q = 0; q < N; ++q)
                if (b.preorder[q] == r) first = q;
            b.b_pre0[nb] = (signed char)first;
            b.b_pre1[nb] = (signed char)(first + cnt);
            for (int q = k; q <= k1; ++q) {
                const int jj = d.ix.dfs_event[q];
                if (jj >= 0 && b.nchild[jj] >= 2) { b.xoff[jj] = (short)(22 * depth + 39 * extra); ++extra; }
            }
            b.b_depth[nb] = (signed char)depth;
            area = std::max(area, 22 * depth + 39 * extra);
            ++nb;
            k = k1 + 1;
        }
        b.num_branches = nb;
        b.branch_area = area;
    }
    {
        int sb = 0, sc = 0;
        for (int j = 0; j < N; ++j) {
            d.ix.slot_base[j] = (j >= 1 && d.ix.valid_body[j]) ? sb++ : -1;
            d.ix.slot_ct[j] = d.ix.valid_body[j] ? sc++ : -1;
        }
    }

    for (int k = 0; k < 3; ++k) d.gravity[k] = m->gravity[k];
    d.num_frames = m->num_frames;
    d.loop = m->loop;
    if (m->num_frames > 0) {
        const int fs = m->n + 1;
        d.motion_duration = m->frames[(size_t)(m->num_frames - 1) * fs];  // cMotion::GetDuration, Motion.cpp:219-224
        for (int k = 0; k < 3; ++k)
            d.cycle_delta[k] = m->frames[(size_t)(m->num_frames - 1) * fs + 1 + k] - m->frames[1 + k];
    }
    return TRL_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: model
// ------------------------------------------------------------------------------------------------
extern "C" trl_model* trl_model_create(int num_joints, const double* joint_mat, const double* body_defs,
                                       const double* pd_params, int num_frames, const double* frames, int loop,
                                       const double* gravity) {
    if (num_joints < 1 || !joint_mat || !body_defs || !pd_params) {
        trl_set_error("trl_model_create: null matrix or num_joints < 1");
        return nullptr;
    }
    std::unique_ptr<trl_model> m(new trl_model());
    m->N = num_joints;
    m->joint_mat.assign(joint_mat, joint_mat + (size_t)num_joints * TRL_JOINT_COLS);
    m->body_defs.assign(body_defs, body_defs + (size_t)num_joints * TRL_BODY_COLS);
    m->pd_params.assign(pd_params, pd_params + (size_t)num_joints * TRL_PD_COLS);
    if (gravity) for (int k = 0; k < 3; ++k) m->gravity[k] = gravity[k];
    // PostProcessJointMat (KinTree.cpp:1026-1041): recompute param offsets, zero the root attach point
    {
        int off = 0;
        for (int j = 0; j < num_joints; ++j) {
            double* jr = &m->joint_mat[(size_t)j * TRL_JOINT_COLS];
            int size = ((int)jr[1] == -1) ? 7 : joint_param_size((int)jr[0]);
            if (size < 0) { trl_set_error("unsupported joint type"); return nullptr; }
            jr[18] = off;
            off += size;
        }
        m->joint_mat[2] = m->joint_mat[3] = m->joint_mat[4] = 0;
        m->n = off;
    }
    m->num_frames = 0;
    m->loop = loop ? 1 : 0;
    if (frames && num_frames > 0) {
        const int fs = m->n + 1;
        m->frames.assign(frames, frames + (size_t)num_frames * fs);
        m->num_frames = num_frames;
        double t = 0;  // cMotion::PostProcessFrames, Motion.cpp:207-217
        for (int f = 0; f < num_frames; ++f) {
            double dur = m->frames[(size_t)f * fs];
            m->frames[(size_t)f * fs] = t;
            t += dur;
        }
    }
    if (trl_build_dev_model(m.get()) != TRL_OK) return nullptr;
    return m.release();
}

extern "C" trl_model* trl_model_load(const char* character_json_path, const char* motion_json_path,
                                     const double* gravity) {
    if (!character_json_path) { trl_set_error("trl_model_load: null path"); return nullptr; }
    std::string text, err;
    if (!read_file(character_json_path, text)) {
        trl_set_error(std::string("Failed to open ") + character_json_path);
        return nullptr;
    }
    JVal root;
    if (!JParser(text).parse(root, err)) {
        trl_set_error(std::string("Failed to parse Json from ") + character_json_path + ": " + err);
        return nullptr;
    }
    const JVal* skel = root.get("Skeleton");  // cCharacter::LoadSkeleton -> cKinTree::Load
    const JVal* joints = skel ? skel->get("Joints") : nullptr;
    const JVal* bodies = root.get("BodyDefs");
    const JVal* pds = root.get("PDControllers");
    if (!joints || joints->kind != JVal::ARR || joints->arr.empty()) {
        trl_set_error("character file has no Skeleton.Joints");
        return nullptr;
    }
    const int N = (int)joints->arr.size();
    std::vector<double> jm((size_t)N * TRL_JOINT_COLS), bd((size_t)N * TRL_BODY_COLS, 0.0), pd((size_t)N * TRL_PD_COLS, 0.0);
    const double inf = std::numeric_limits<double>::infinity();
    for (int j = 0; j < N; ++j) {
        // cKinTree::BuildJointDesc defaults (KinTree.cpp:1153-1177) then ParseJoint (:980-1008)
        double* r = &jm[(size_t)j * TRL_JOINT_COLS];
        const double def[TRL_JOINT_COLS] = {0, -1, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0, inf, inf, 0, 1, 0};
        std::memcpy(r, def, sizeof(def));
        const JVal& jj = joints->arr[j];
        const JVal* t = jj.get("Type");
        int type = JT_NONE;
        if (t && t->kind == JVal::STR) {
            bool found = false;
            for (int k = 0; k < 6; ++k)
                if (t->str == kJointTypeNames[k]) { type = k; found = true; }
            if (!found) { trl_set_error("Unsupported joint type: " + t->str); return nullptr; }
        }
        r[0] = type;
        for (int k = 1; k < TRL_JOINT_COLS; ++k) {
            const JVal* v = jj.get(kJointKeys[k]);
            if (v && !v->is_null()) r[k] = v->as_double();
        }
    }
    if (bodies && bodies->kind == JVal::ARR) {
        if ((int)bodies->arr.size() != N) { trl_set_error("BodyDefs count != joint count"); return nullptr; }
        for (int j = 0; j < N; ++j) {
            // BuildBodyDef defaults (KinTree.cpp:1179-1200) then ParseBodyDef (:176-199)
            double* r = &bd[(size_t)j * TRL_BODY_COLS];
            r[2] = -1;
            r[16] = 1;
            const JVal& bj = bodies->arr[j];
            const JVal* s = bj.get("Shape");
            std::string shape = (s && s->kind == JVal::STR) ? s->str : "";
            if (shape == "box") r[0] = 0;
            else if (shape == "capsule") r[0] = 1;
            else if (shape == "null") r[0] = 3;
            else { trl_set_error("Unsupported body shape " + shape); return nullptr; }
            for (int k = 0; k < TRL_BODY_COLS; ++k) {
                const JVal* v = bj.get(kBodyKeys[k]);
                if (v && !v->is_null() && v->is_numeric()) r[k] = v->as_double();
            }
        }
    } else {
        trl_set_error("Failed to load body definition");
        return nullptr;
    }
    if (pds && pds->kind == JVal::ARR) {
        if ((int)pds->arr.size() != N) { trl_set_error("PDControllers count != joint count"); return nullptr; }
        for (int j = 0; j < N; ++j) {  // cPDController::ParsePDParams (PDController.cpp:73-91), JointID = index (:55)
            double* r = &pd[(size_t)j * TRL_PD_COLS];
            for (int k = 0; k < TRL_PD_COLS; ++k) {
                const JVal* v = pds->arr[j].get(kPDKeys[k]);
                if (v && !v->is_null() && v->is_numeric()) r[k] = v->as_double();
            }
            r[0] = j;
        }
    } else {
        for (int j = 0; j < N; ++j) pd[(size_t)j * TRL_PD_COLS] = j;
    }

    std::vector<double> frames;
    int num_frames = 0, loop = 0;
    if (motion_json_path && motion_json_path[0]) {
        std::string mt;
        if (!read_file(motion_json_path, mt)) {
            trl_set_error(std::string("Failed to load motion from file ") + motion_json_path);
            return nullptr;
        }
        JVal mr;
        if (!JParser(mt).parse(mr, err)) {
            trl_set_error(std::string("Failed to parse Json from ") + motion_json_path + ": " + err);
            return nullptr;
        }
        const JVal* lp = mr.get("Loop");
        if (lp && !lp->is_null()) loop = lp->as_double() != 0;
        const JVal* fr = mr.get("Frames");
        if (fr && fr->kind == JVal::ARR && !fr->arr.empty()) {
            num_frames = (int)fr->arr.size();
            size_t fs = fr->arr[0].arr.size();
            for (int f = 0; f < num_frames; ++f) {
                if (fr->arr[f].kind != JVal::ARR || fr->arr[f].arr.size() != fs) {
                    trl_set_error("motion frames have inconsistent sizes");
                    return nullptr;
                }
                for (size_t i = 0; i < fs; ++i) frames.push_back(fr->arr[f].arr[i].as_double());
            }
            // dof check of cKinCharacter::LoadMotion (anim/KinCharacter.cpp:205-218)
            int n = 0;
            for (int j = 0; j < N; ++j) n += ((int)jm[(size_t)j * TRL_JOINT_COLS + 1] == -1) ? 7 : std::max(0, joint_param_size((int)jm[(size_t)j * TRL_JOINT_COLS]));
            if ((int)fs != n + 1) {
                char buf[128];
                std::snprintf(buf, sizeof(buf), "DOF mismatch, char dof: %i, motion dof: %i", n, (int)fs - 1);
                trl_set_error(buf);
                return nullptr;
            }
        }
    }
    return trl_model_create(N, jm.data(), bd.data(), pd.data(), num_frames, frames.empty() ? nullptr : frames.data(),
                            loop, gravity);
}

extern "C" int trl_model_dims(const trl_model* m, int* num_joints, int* num_dof, int* num_frames) {
    if (!m) return TRL_ERR_INVALID_ARG;
    if (num_joints) *num_joints = m->N;
    if (num_dof) *num_dof = m->n;
    if (num_frames) *num_frames = m->num_frames;
    return TRL_OK;
}

extern "C" int trl_model_get_matrices(const trl_model* m, double* joint_mat, double* body_defs, double* pd_params) {
    if (!m) return TRL_ERR_INVALID_ARG;
    if (joint_mat) std::memcpy(joint_mat, m->joint_mat.data(), m->joint_mat.size() * sizeof(double));
    if (body_defs) std::memcpy(body_defs, m->body_defs.data(), m->body_defs.size() * sizeof(double));
    if (pd_params) std::memcpy(pd_params, m->pd_params.data(), m->pd_params.size() * sizeof(double));
    return TRL_OK;
}

extern "C" int trl_model_get_motion(const trl_model* m, double* frames, int* loop) {
    if (!m) return TRL_ERR_INVALID_ARG;
    if (m->num_frames == 0) return TRL_ERR_NO_MOTION;
    if (frames) std::memcpy(frames, m->frames.data(), m->frames.size() * sizeof(double));
    if (loop) *loop = m->loop;
    return TRL_OK;
}

extern "C" void trl_model_free(trl_model* m) { delete m; }

// cArgParser file format (util/ArgParser.cpp): whitespace-separated tokens, "-key=" starts a key, the tokens
// that follow (until the next key) are its values; "//" starts a comment to end of line.
extern "C" int trl_arg_file_get(const char* path, const char* key, char* out_value, int out_capacity) {
    if (!path || !key || !out_value || out_capacity < 1) return TRL_ERR_INVALID_ARG;
    std::string text;
    if (!read_file(path, text)) { trl_set_error(std::string("Failed to open ") + path); return TRL_ERR_IO; }
    std::istringstream lines(text);
    std::string line, cur_key, found;
    bool have = false;
    while (std::getline(lines, line)) {
        size_t c = line.find("//");
        if (c != std::string::npos) line = line.substr(0, c);
        std::istringstream toks(line);
        std::string tok;
        while (toks >> tok) {
            if (tok.size() > 2 && tok[0] == '-' && tok.back() == '=' && !(tok.size() > 1 && (isdigit(tok[1]) || tok[1] == '.'))) {
                cur_key = tok.substr(1, tok.size() - 2);
                if (cur_key == key) { have = true; found.clear(); }
            } else if (cur_key == key && have) {
                if (!found.empty()) found += ' ';
                found += tok;
            }
        }
    }
    if (!have) return TRL_ERR_INVALID_ARG;
    std::snprintf(out_value, (size_t)out_capacity, "%s", found.c_str());
    return TRL_OK;
}

// ------------------------------------------------------------------------------------------------
// terrain parameter files (data/terrain/*.txt): cGroundFactory::ParseParamsJson (sim/GroundFactory.cpp:11-60),
// cTerrainGen2D::LoadParams (sim/TerrainGen2D.cpp:72-84), cGround::CalcBlendParams (sim/Ground.cpp:255-277)
// ------------------------------------------------------------------------------------------------
namespace {
struct TerrainParamDef { const char* name; double dflt; };
const TerrainParamDef kTerrainParams[TRL_TERRAIN_NUM_PARAMS] = {  // cTerrainGen2D::gParamDefs, sim/TerrainGen2D.cpp:12-63
    {"GapSpacingMin", 4}, {"GapSpacingMax", 7}, {"GapWMin", 0.5}, {"GapWMax", 2}, {"GapHMin", -2}, {"GapHMax", -2},
    {"WallSpacingMin", 6}, {"WallSpacingMax", 8}, {"WallWMin", 0.2}, {"WallWMax", 0.2}, {"WallHMin", 0.25}, {"WallHMax", 0.5},
    {"StepSpacingMin", 5}, {"StepSpacingMax", 7}, {"StepH0Min", 0.1}, {"StepH0Max", 0.4}, {"StepH1Min", -0.4}, {"StepH1Max", -0.1},
    {"BumpHMin", 0}, {"BumpHMax", 0.03},
    {"NarrowGapSpacingMin", 3}, {"NarrowGapSpacingMax", 6}, {"NarrowGapDistMin", 0.1}, {"NarrowGapDistMax", 0.4},
    {"NarrowGapWMin", 0.15}, {"NarrowGapWMax", 0.5}, {"NarrowGapDepthMin", -2}, {"NarrowGapDepthMax", -2},
    {"NarrowGapCountMin", 1}, {"NarrowGapCountMax", 4},
    {"CliffSpacingMin", 5}, {"CliffSpacingMax", 7}, {"CliffH0Min", 0.1}, {"CliffH0Max", 0.4}, {"CliffH1Min", -0.4},
    {"CliffH1Max", -0.1}, {"CliffMiniCountMax", 0}, {"SlopeDeltaRange", 0.25}, {"SlopeDeltaMin", -0.35}, {"SlopeDeltaMax", 0.35}};
const char* const kTerrainTypeNames[] = {  // cGround::gGroundTypeNames (sim/Ground.cpp:5-33), the types this library builds
    "plane", "var2d_flat", "var2d_gaps", "var2d_steps", "var2d_walls", "var2d_bumps", "var2d_mixed", "var2d_narrow_gaps",
    "var2d_slopes", "var2d_slopes_gaps", "var2d_slopes_walls", "var2d_slopes_steps", "var2d_slopes_mixed",
    "var2d_slopes_narrow_gaps", "var2d_cliffs", "var3d_flat"};
}  // namespace

extern "C" int trl_terrain_default_cfg(trl_terrain_cfg* out) {
    if (!out) return TRL_ERR_INVALID_ARG;
    out->type = TRL_TERRAIN_VAR2D_FLAT;  // cGround::tParams::tParams, sim/Ground.cpp:41-54
    out->pad = 0;
    for (int i = 0; i < TRL_TERRAIN_NUM_PARAMS; ++i) out->params[i] = kTerrainParams[i].dflt;
    out->ground_width = 20;
    out->world_scale = 4;
    out->spacing_x = out->spacing_z = 0.2;
    return TRL_OK;
}

extern "C" int trl_terrain_load(const char* terrain_file, double blend, trl_terrain_cfg* out) {
    if (!terrain_file || !out) return TRL_ERR_INVALID_ARG;
    trl_terrain_default_cfg(out);
    std::string text, err;
    if (!read_file(terrain_file, text)) { trl_set_error(std::string("Failed to open ") + terrain_file); return TRL_ERR_IO; }
    JVal root;
    JParser parser(text);
    if (!parser.parse(root, err)) { trl_set_error(std::string("Failed to parse ") + terrain_file + ": " + err); return TRL_ERR_PARSE; }
    if (const JVal* t = root.get("Type")) {
        int found = -1;
        for (int i = 0; i < (int)(sizeof(kTerrainTypeNames) / sizeof(kTerrainTypeNames[0])); ++i)
            if (t->str == kTerrainTypeNames[i]) found = i;
        if (found < 0) {
            trl_set_error("trl_terrain_load: ground type '" + t->str + "' is not built by this library");
            return TRL_ERR_UNSUPPORTED;
        }
        out->type = found;
    }
    if (const JVal* v = root.get("GroundWidth")) out->ground_width = v->as_double();
    if (const JVal* v = root.get("VertSpacingX")) out->spacing_x = v->as_double();
    if (const JVal* v = root.get("VertSpacingZ")) out->spacing_z = v->as_double();
    const JVal* arr = root.get("Params");
    if (arr && arr->kind == JVal::ARR && !arr->arr.empty() && out->type != TRL_TERRAIN_PLANE && out->type != TRL_TERRAIN_VAR3D_FLAT) {
        const int num = (int)arr->arr.size();
        std::vector<double> rows((size_t)num * TRL_TERRAIN_NUM_PARAMS);
        for (int r = 0; r < num; ++r)
            for (int i = 0; i < TRL_TERRAIN_NUM_PARAMS; ++i) {
                const JVal* v = arr->arr[r].get(kTerrainParams[i].name);
                rows[(size_t)r * TRL_TERRAIN_NUM_PARAMS + i] = (v && !v->is_null()) ? v->as_double() : kTerrainParams[i].dflt;
            }
        double bl = std::max(0.0, std::min(blend, num - 1.0));  // cMathUtil::Clamp
        const int idx0 = (int)bl, idx1 = std::min(idx0 + 1, num - 1);
        bl -= idx0;
        for (int i = 0; i < TRL_TERRAIN_NUM_PARAMS; ++i)
            out->params[i] = (1 - bl) * rows[(size_t)idx0 * TRL_TERRAIN_NUM_PARAMS + i] + bl * rows[(size_t)idx1 * TRL_TERRAIN_NUM_PARAMS + i];
    }
    return TRL_OK;
}

// ------------------------------------------------------------------------------------------------
// character state files (data/states/*.txt): cCharacter::ReadState / WriteState / BuildStateJson
// (anim/Character.cpp:277-342,362-379), cJsonUtil::BuildVectorJson (util/JsonUtil.cpp:40-59: std::to_string, "%f")
// ------------------------------------------------------------------------------------------------
extern "C" int trl_state_read(const trl_model* m, const char* state_file, double* inout_pose, double* inout_vel) {
    if (!m || !state_file) return TRL_ERR_INVALID_ARG;
    std::string text, err;
    if (!read_file(state_file, text)) { trl_set_error(std::string("Failed to open ") + state_file); return TRL_ERR_IO; }
    JVal root;
    JParser parser(text);
    if (!parser.parse(root, err)) { trl_set_error(std::string("Failed to parse ") + state_file + ": " + err); return TRL_ERR_PARSE; }
    const int n = m->n, N = m->N;
    auto read_vec = [&](const char* key, double* out) -> int {
        const JVal* v = root.get(key);
        if (!v || v->is_null() || !out) return TRL_OK;  // a missing key keeps the caller's values (ReadState :326,334)
        if (v->kind != JVal::ARR || (int)v->arr.size() != n) {
            trl_set_error(std::string(state_file) + ": \"" + key + "\" must hold " + std::to_string(n) + " numbers");
            return TRL_ERR_PARSE;
        }
        for (int i = 0; i < n; ++i) out[i] = v->arr[i].as_double();
        return TRL_OK;
    };
    int rc = read_vec("Pose", inout_pose);
    if (rc == TRL_OK && inout_pose && root.get("Pose") && !root.get("Pose")->is_null()) {
        // cKinTree::PoseProcessPose (anim/KinTree.cpp:1510-1527): normalise the root and the spherical joints' quaternions
        auto normalize4 = [](double* q) {
            const double nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            if (nrm > 0) for (int k = 0; k < 4; ++k) q[k] /= nrm;
        };
        normalize4(inout_pose + 3);
        for (int j = 1; j < N; ++j)
            if ((int)m->joint_mat[(size_t)j * 19] == JT_SPHERICAL) normalize4(inout_pose + (int)m->joint_mat[(size_t)j * 19 + 18]);
    }
    if (rc == TRL_OK) rc = read_vec("Vel", inout_vel);
    return rc;
}

extern "C" int trl_state_write(const trl_model* m, const char* state_file, const double* pose, const double* vel) {
    if (!m || !state_file || !pose || !vel) return TRL_ERR_INVALID_ARG;
    auto vec_json = [&](const double* v) {
        std::string out = "[";
        for (int i = 0; i < m->n; ++i) {
            if (i) out += ", ";
            out += std::to_string(v[i]);
        }
        return out + "]";
    };
    const std::string json = "{\n\"Pose\":" + vec_json(pose) + ",\n\"Vel\":" + vec_json(vel) + "\n}";
    FILE* f = std::fopen(state_file, "w");
    if (!f) { trl_set_error(std::string("Failed to open ") + state_file); return TRL_ERR_IO; }
    std::fprintf(f, "%s", json.c_str());
    std::fclose(f);
    return TRL_OK;
}
