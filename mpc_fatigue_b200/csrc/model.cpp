// model.cpp — URDF reader (no urdfdom / Eigen in this image: a small XML scanner is enough for URDF)
// and synthetic tree generators.  See model.hpp for what it replaces in the reference.
#include "model.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>

namespace mpcf {

namespace {

// ---------- 3x3 helpers (row-major) ----------
void mat_mul(const double A[9], const double B[9], double C[9])
{
    double t[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) t[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
    std::memcpy(C, t, sizeof t);
}
void mat_mul_t(const double A[9], const double B[9], double C[9])  // A * B^T
{
    double t[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
            t[3 * r + c] = A[3 * r] * B[3 * c] + A[3 * r + 1] * B[3 * c + 1] + A[3 * r + 2] * B[3 * c + 2];
    std::memcpy(C, t, sizeof t);
}
void mat_vec(const double A[9], const double v[3], double o[3])
{
    double t[3] = {A[0] * v[0] + A[1] * v[1] + A[2] * v[2], A[3] * v[0] + A[4] * v[1] + A[5] * v[2],
                   A[6] * v[0] + A[7] * v[1] + A[8] * v[2]};
    std::memcpy(o, t, sizeof t);
}
void transpose(const double A[9], double T[9])
{
    double t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
    std::memcpy(T, t, sizeof t);
}
const double kEye[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};

// URDF rpy: R = Rz(yaw) * Ry(pitch) * Rx(roll)
void rpy_to_R(double r, double p, double y, double R[9])
{
    double cr = std::cos(r), sr = std::sin(r), cp = std::cos(p), sp = std::sin(p), cy = std::cos(y), sy = std::sin(y);
    double Rx[9] = {1, 0, 0, 0, cr, -sr, 0, sr, cr};
    double Ry[9] = {cp, 0, sp, 0, 1, 0, -sp, 0, cp};
    double Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1};
    double t[9];
    mat_mul(Rz, Ry, t);
    mat_mul(t, Rx, R);
}

// rotation with R z = a (Rodrigues about z x a); identity when a = +z
void axis_to_R(const double a_in[3], double R[9])
{
    double nrm = std::sqrt(a_in[0] * a_in[0] + a_in[1] * a_in[1] + a_in[2] * a_in[2]);
    double a[3] = {a_in[0] / nrm, a_in[1] / nrm, a_in[2] / nrm};
    if (std::fabs(a[0]) < 1e-14 && std::fabs(a[1]) < 1e-14 && a[2] > 0) { std::memcpy(R, kEye, sizeof kEye); return; }
    if (std::fabs(a[0]) < 1e-14 && std::fabs(a[1]) < 1e-14 && a[2] < 0) {
        double t[9] = {1, 0, 0, 0, -1, 0, 0, 0, -1};
        std::memcpy(R, t, sizeof t);
        return;
    }
    double v[3] = {-a[1], a[0], 0.0};  // z x a
    double c = a[2];
    double vx[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
    double vx2[9];
    mat_mul(vx, vx, vx2);
    for (int k = 0; k < 9; ++k) R[k] = kEye[k] + vx[k] + vx2[k] / (1.0 + c);
}

// ---------- minimal XML ----------
struct XmlNode {
    std::string name;
    std::map<std::string, std::string> attr;
    std::vector<std::unique_ptr<XmlNode>> kids;
    const XmlNode *child(const char *n) const
    {
        for (auto &k : kids)
            if (k->name == n) return k.get();
        return nullptr;
    }
    const std::string *get(const char *k) const
    {
        auto it = attr.find(k);
        return it == attr.end() ? nullptr : &it->second;
    }
};

struct XmlParser {
    const char *s, *e;
    std::string err;
    bool fail(const std::string &m) { if (err.empty()) err = m; return false; }
    void skip_ws() { while (s < e && std::isspace((unsigned char)*s)) ++s; }
    bool starts(const char *lit) const { size_t n = std::strlen(lit); return (size_t)(e - s) >= n && std::memcmp(s, lit, n) == 0; }
    bool skip_until(const char *lit)
    {
        size_t n = std::strlen(lit);
        while ((size_t)(e - s) >= n) {
            if (std::memcmp(s, lit, n) == 0) { s += n; return true; }
            ++s;
        }
        return fail(std::string("unterminated construct, expected ") + lit);
    }
    // skip comments, processing instructions, doctype and text; stop at '<' of an element / closing tag
    bool skip_misc()
    {
        for (;;) {
            while (s < e && *s != '<') ++s;
            if (s >= e) return true;
            if (starts("<!--")) { if (!skip_until("-->")) return false; }
            else if (starts("<?")) { if (!skip_until("?>")) return false; }
            else if (starts("<![CDATA[")) { if (!skip_until("]]>")) return false; }
            else if (starts("<!")) { if (!skip_until(">")) return false; }
            else return true;
        }
    }
    static bool name_char(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-' || c == ':' || c == '.'; }
    bool parse_element(XmlNode &node, int depth)
    {
        if (depth > 64) return fail("XML nesting too deep");
        if (s >= e || *s != '<') return fail("expected '<'");
        ++s;
        const char *b = s;
        while (s < e && name_char(*s)) ++s;
        if (s == b) return fail("empty element name");
        node.name.assign(b, s);
        for (;;) {
            skip_ws();
            if (s >= e) return fail("unexpected end inside tag <" + node.name + ">");
            if (*s == '/') {
                if (s + 1 < e && s[1] == '>') { s += 2; return true; }
                return fail("stray '/' in tag <" + node.name + ">");
            }
            if (*s == '>') { ++s; break; }
            const char *ab = s;
            while (s < e && name_char(*s)) ++s;
            if (s == ab) return fail("bad attribute in <" + node.name + ">");
            std::string key(ab, s);
            skip_ws();
            if (s >= e || *s != '=') return fail("attribute '" + key + "' without value");
            ++s;
            skip_ws();
            if (s >= e || (*s != '"' && *s != '\'')) return fail("attribute '" + key + "' value not quoted");
            char qc = *s++;
            const char *vb = s;
            while (s < e && *s != qc) ++s;
            if (s >= e) return fail("unterminated attribute value");
            node.attr[key] = std::string(vb, s);
            ++s;
        }
        for (;;) {  // children
            if (!skip_misc()) return false;
            if (s >= e) return fail("missing </" + node.name + ">");
            if (s + 1 < e && s[1] == '/') {
                s += 2;
                const char *cb = s;
                while (s < e && name_char(*s)) ++s;
                if (std::string(cb, s) != node.name) return fail("mismatched closing tag </" + std::string(cb, s) + "> for <" + node.name + ">");
                skip_ws();
                if (s >= e || *s != '>') return fail("bad closing tag");
                ++s;
                return true;
            }
            node.kids.emplace_back(new XmlNode());
            if (!parse_element(*node.kids.back(), depth + 1)) return false;
        }
    }
};

bool parse_floats(const std::string *s, int n, double *out, double dflt)
{
    for (int i = 0; i < n; ++i) out[i] = dflt;
    if (!s) return true;
    const char *p = s->c_str();
    for (int i = 0; i < n; ++i) {
        char *end = nullptr;
        out[i] = std::strtod(p, &end);
        if (end == p) return false;
        p = end;
    }
    return true;
}
bool parse_float(const std::string *s, double *out, double dflt)
{
    *out = dflt;
    if (!s) return true;
    char *end = nullptr;
    *out = std::strtod(s->c_str(), &end);
    return end != s->c_str();
}

struct LinkBody { bool has = false; double m = 0, c[3] = {0, 0, 0}, Ic[9] = {0}; };

}  // namespace

// ---------- HostModel ----------
int HostModel::add_joint(int parent_joint, int type, const std::string &name, const double R[9], const double p[3],
                         const mpcf_opts &o)
{
    int idx = n++;
    parent.push_back(parent_joint);
    jtype.push_back(type);
    jcontinuous.push_back(0);
    joint_names.push_back(name);
    Rp.insert(Rp.end(), R, R + 9);
    pp.insert(pp.end(), p, p + 3);
    mass.push_back(0.0);
    mc.insert(mc.end(), 3, 0.0);
    Io.insert(Io.end(), 6, 0.0);
    arm.push_back(o.armature);
    const double row[4] = {o.lambda, o.kappa, o.ctau, o.cv};
    fat.insert(fat.end(), row, row + 4);
    q_lo.push_back(-M_PI);
    q_hi.push_back(M_PI);
    v_max.push_back(1.0);
    tau_max.push_back(1.0);
    return idx;
}

void HostModel::add_frame(const std::string &name, int joint, const double R[9], const double p[3])
{
    frame_names.push_back(name);
    fparent.push_back(joint);
    fR.insert(fR.end(), R, R + 9);
    fp.insert(fp.end(), p, p + 3);
}

void HostModel::add_body(int j, double m, const double c[3], const double Ic[9], const double R[9], const double p[3])
{
    if (j < 0) return;  // fixed to the world: no dynamics
    double cj[3], t[9], Icj[9];
    mat_vec(R, c, cj);
    for (int k = 0; k < 3; ++k) cj[k] += p[k];
    mat_mul(R, Ic, t);
    mat_mul_t(t, R, Icj);
    double cc = cj[0] * cj[0] + cj[1] * cj[1] + cj[2] * cj[2];
    double I[9];
    for (int r = 0; r < 3; ++r)
        for (int q = 0; q < 3; ++q) I[3 * r + q] = Icj[3 * r + q] + m * ((r == q ? cc : 0.0) - cj[r] * cj[q]);
    mass[j] += m;
    for (int k = 0; k < 3; ++k) mc[3 * j + k] += m * cj[k];
    double *o = &Io[6 * j];
    o[0] += I[0]; o[1] += I[1]; o[2] += I[2]; o[3] += I[4]; o[4] += I[5]; o[5] += I[8];
}

// ---------- URDF ----------
namespace {

struct UrdfCtx {
    const mpcf_opts *opts;
    HostModel *mdl;
    std::map<std::string, LinkBody> links;
    std::map<std::string, std::vector<const XmlNode *>> joints_of;  // parent link -> joints in document order
    std::string err;
    int code = MPCF_OK;

    bool visit(const std::string &link, int jidx, const double R[9], const double p[3], int depth)
    {
        if (depth > 4096) { err = "URDF kinematic loop"; code = MPCF_EPARSE; return false; }
        mdl->add_frame(link, jidx, R, p);
        const LinkBody &b = links[link];
        if (b.has) mdl->add_body(jidx, b.m, b.c, b.Ic, R, p);
        for (const XmlNode *je : joints_of[link]) {
            const XmlNode *o = je->child("origin");
            double xyz[3], rpy[3];
            if (!parse_floats(o ? o->get("xyz") : nullptr, 3, xyz, 0.0) || !parse_floats(o ? o->get("rpy") : nullptr, 3, rpy, 0.0)) {
                err = "bad <origin> in joint " + *je->get("name"); code = MPCF_EPARSE; return false;
            }
            double Ro[9], Rj[9], pj[3];
            rpy_to_R(rpy[0], rpy[1], rpy[2], Ro);
            mat_mul(R, Ro, Rj);
            mat_vec(R, xyz, pj);
            for (int k = 0; k < 3; ++k) pj[k] += p[k];
            const std::string &type = *je->get("type");
            const std::string &jname = *je->get("name");
            const std::string &child = *je->child("child")->get("link");
            if (type == "fixed") {
                mdl->add_frame(jname, jidx, Rj, pj);
                if (!visit(child, jidx, Rj, pj, depth + 1)) return false;
                continue;
            }
            int jt;
            if (type == "revolute" || type == "continuous") jt = 0;
            else if (type == "prismatic") jt = 1;
            else { err = "unsupported joint type '" + type + "' (joint " + jname + ")"; code = MPCF_EJOINT; return false; }
            if (mdl->n >= MPCF_MAX_DOF) { err = "model has more than MPCF_MAX_DOF joints"; code = MPCF_ELIMIT; return false; }
            const XmlNode *ax = je->child("axis");
            double a[3];
            if (!parse_floats(ax ? ax->get("xyz") : nullptr, 3, a, 0.0)) { err = "bad <axis> in joint " + jname; code = MPCF_EPARSE; return false; }
            if (!ax || !ax->get("xyz")) { a[0] = 1; a[1] = 0; a[2] = 0; }  // URDF default axis (element or attribute missing)
            if (a[0] == 0 && a[1] == 0 && a[2] == 0) { err = "zero <axis> in joint " + jname; code = MPCF_EPARSE; return false; }
            double Ra[9], Rn[9], RaT[9];
            axis_to_R(a, Ra);
            mat_mul(Rj, Ra, Rn);
            int idx = mdl->add_joint(jidx, jt, jname, Rn, pj, *opts);
            mdl->jcontinuous[idx] = type == "continuous";
            if (const XmlNode *lim = je->child("limit")) {
                parse_float(lim->get("lower"), &mdl->q_lo[idx], -M_PI);
                parse_float(lim->get("upper"), &mdl->q_hi[idx], M_PI);
                parse_float(lim->get("velocity"), &mdl->v_max[idx], 1.0);
                parse_float(lim->get("effort"), &mdl->tau_max[idx], 1.0);
            }
            const double zero[3] = {0, 0, 0};
            // The joint's own frame keeps the URDF (un-normalised) orientation, as Pinocchio's JOINT frame does: its pose is
            // Rj * Rot(axis, q) = (Rj Ra) Rz(q) Ra^T, i.e. Ra^T relative to the axis-normalised joint frame (identity when the
            // axis is +z, as in every shipped Pilz URDF).
            transpose(Ra, RaT);
            mdl->add_frame(jname, idx, RaT, zero);
            if (!visit(child, idx, RaT, zero, depth + 1)) return false;
        }
        return true;
    }
};

}  // namespace

int parse_urdf(const char *xml, size_t len, const mpcf_opts &opts, HostModel &out, std::string &err)
{
    XmlParser xp{xml, xml + len, {}};
    XmlNode root;
    if (!xp.skip_misc() || xp.s >= xp.e) { err = xp.err.empty() ? "empty document" : xp.err; return MPCF_EPARSE; }
    if (!xp.parse_element(root, 0)) { err = xp.err; return MPCF_EPARSE; }
    if (root.name != "robot") { err = "root element is <" + root.name + ">, expected <robot>"; return MPCF_EPARSE; }

    UrdfCtx ctx;
    ctx.opts = &opts;
    ctx.mdl = &out;
    std::vector<std::string> link_order;
    std::map<std::string, bool> is_child;
    for (auto &k : root.kids) {
        if (k->name != "link") continue;
        const std::string *name = k->get("name");
        if (!name) { err = "<link> without name"; return MPCF_EPARSE; }
        LinkBody b;
        if (const XmlNode *in = k->child("inertial")) {
            const XmlNode *o = in->child("origin");
            double rpy[3], I6[6] = {0, 0, 0, 0, 0, 0};
            const XmlNode *me = in->child("mass");
            if (!me || !parse_float(me->get("value"), &b.m, 0.0)) { err = "bad <mass> in link " + *name; return MPCF_EPARSE; }
            if (!parse_floats(o ? o->get("xyz") : nullptr, 3, b.c, 0.0) || !parse_floats(o ? o->get("rpy") : nullptr, 3, rpy, 0.0)) {
                err = "bad inertial <origin> in link " + *name; return MPCF_EPARSE;
            }
            if (const XmlNode *ie = in->child("inertia")) {
                const char *keys[6] = {"ixx", "ixy", "ixz", "iyy", "iyz", "izz"};
                for (int i = 0; i < 6; ++i)
                    if (!parse_float(ie->get(keys[i]), &I6[i], 0.0)) { err = "bad <inertia> in link " + *name; return MPCF_EPARSE; }
            }
            double I[9] = {I6[0], I6[1], I6[2], I6[1], I6[3], I6[4], I6[2], I6[4], I6[5]}, Rin[9], t[9];
            rpy_to_R(rpy[0], rpy[1], rpy[2], Rin);
            mat_mul(Rin, I, t);
            mat_mul_t(t, Rin, b.Ic);
            b.has = true;
        }
        if (ctx.links.count(*name)) { err = "duplicate link " + *name; return MPCF_EPARSE; }
        ctx.links[*name] = b;
        link_order.push_back(*name);
    }
    for (auto &k : root.kids) {
        if (k->name != "joint") continue;
        const std::string *name = k->get("name"), *type = k->get("type");
        const XmlNode *pe = k->child("parent"), *ce = k->child("child");
        if (!name || !type || !pe || !ce || !pe->get("link") || !ce->get("link")) { err = "incomplete <joint>"; return MPCF_EPARSE; }
        const std::string &par = *pe->get("link"), &ch = *ce->get("link");
        if (!ctx.links.count(par) || !ctx.links.count(ch)) { err = "joint " + *name + " references an unknown link"; return MPCF_EPARSE; }
        if (is_child[ch]) { err = "link " + ch + " has two parents"; return MPCF_EPARSE; }
        is_child[ch] = true;
        ctx.joints_of[par].push_back(k.get());
    }
    std::string root_link;
    int nroots = 0;
    for (auto &l : link_order)
        if (!is_child[l]) { root_link = l; ++nroots; }
    if (nroots != 1) { err = "URDF must have exactly one root link"; return MPCF_EPARSE; }
    const double zero[3] = {0, 0, 0};
    std::memcpy(out.grav, opts.gravity, sizeof out.grav);
    if (!ctx.visit(root_link, -1, kEye, zero, 0)) { err = ctx.err; return ctx.code; }
    if (out.n == 0) { err = "URDF has no moving joints"; return MPCF_EPARSE; }
    return MPCF_OK;
}

// ---------- synthetic trees ----------
namespace {
struct Rng {  // splitmix64
    unsigned long long s;
    unsigned long long next() { unsigned long long z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double range(double a, double b) { return a + (b - a) * uni(); }
    double logu(double a, double b) { return std::exp(range(std::log(a), std::log(b))); }
};

// one synthetic limb link: revolute joint about a pseudo-random axis, Pilz-magnitude inertials
int add_random_link(HostModel &m, Rng &r, int parent, const std::string &name, const mpcf_opts &o, double reach, double mscale)
{
    double rpy[3] = {r.range(-M_PI, M_PI), r.range(-M_PI, M_PI), r.range(-M_PI, M_PI)};
    // snap two of three angles to multiples of pi/2 so the tree looks like a real robot (axes mostly orthogonal)
    rpy[0] = std::round(rpy[0] / (M_PI / 2)) * (M_PI / 2);
    rpy[2] = std::round(rpy[2] / (M_PI / 2)) * (M_PI / 2);
    double R[9], p[3] = {r.range(-0.05, 0.05), r.range(-0.05, 0.05), r.range(0.3, 1.0) * reach};
    rpy_to_R(rpy[0], rpy[1], rpy[2], R);
    int j = m.add_joint(parent, 0, name, R, p, o);
    double ms = r.logu(0.5, 5.0) * mscale;
    double c[3] = {r.range(-0.03, 0.03), r.range(-0.06, 0.06), r.range(0.02, 0.2)};
    double d[3] = {r.logu(3e-3, 4e-2) * mscale, r.logu(3e-3, 4e-2) * mscale, r.logu(3e-3, 4e-2) * mscale};
    // keep the triangle inequality of principal moments
    double mx = std::max(d[0], std::max(d[1], d[2]));
    for (double &v : d) v = std::max(v, 0.55 * mx);
    double Ic[9] = {d[0], 0, 0, 0, d[1], 0, 0, 0, d[2]};
    double Rb[9];
    rpy_to_R(r.range(-0.3, 0.3), r.range(-0.3, 0.3), r.range(-0.3, 0.3), Rb);
    const double zero[3] = {0, 0, 0};
    m.add_body(j, ms, c, Ic, Rb, zero);
    m.q_lo[j] = -2.5; m.q_hi[j] = 2.5; m.v_max[j] = 1.57; m.tau_max[j] = 150.0 * mscale;
    m.add_frame(name, j, kEye, zero);
    m.add_frame(name + "_link", j, kEye, zero);
    return j;
}
}  // namespace

int make_synthetic(int kind, int ndof, unsigned long long seed, const mpcf_opts &opts, HostModel &m, std::string &err)
{
    if (ndof <= 0 || ndof > MPCF_MAX_DOF) { err = "ndof out of range"; return MPCF_ELIMIT; }
    Rng r{seed ? seed : 1};
    std::memcpy(m.grav, opts.gravity, sizeof m.grav);
    const double zero[3] = {0, 0, 0};
    m.add_frame("world", -1, kEye, zero);
    if (kind == MPCF_SYNTH_CHAIN) {
        int par = -1;
        for (int i = 0; i < ndof; ++i) par = add_random_link(m, r, par, "joint_" + std::to_string(i + 1), opts, 0.3, 1.0);
        return MPCF_OK;
    }
    if (kind == MPCF_SYNTH_DUAL_ARM) {  // two serial arms of ndof/2 joints on a fixed torso (the reference's Centauro layout: 2 x 7)
        if (ndof % 2) { err = "dual-arm model needs an even ndof"; return MPCF_EINVAL; }
        const char *side[2] = {"left", "right"};
        for (int a = 0; a < 2; ++a) {
            int par = -1;
            for (int i = 0; i < ndof / 2; ++i) {
                par = add_random_link(m, r, par, std::string(side[a]) + "_joint_" + std::to_string(i + 1), opts, 0.3, 1.0);
                if (i == 0) m.pp[3 * par + 1] += a == 0 ? 0.25 : -0.25;  // shoulders 0.5 m apart
            }
            double pe[3] = {0, 0, 0.1};
            m.add_frame(std::string(side[a]) + "_ee", par, kEye, pe);
        }
        return MPCF_OK;
    }
    if (kind != MPCF_SYNTH_HUMANOID) { err = "unknown synthetic kind"; return MPCF_EINVAL; }
    if (ndof < 8) { err = "humanoid tree needs ndof >= 8"; return MPCF_EINVAL; }
    // floating base as a root chain with nq = nv: prismatic x, y, z then revolute z, y, x.
    const double axes[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 1}, {0, 1, 0}, {1, 0, 0}};
    const char *names[6] = {"root_px", "root_py", "root_pz", "root_rz", "root_ry", "root_rx"};
    int par = -1;
    double Rprev_T[9];
    std::memcpy(Rprev_T, kEye, sizeof kEye);
    for (int i = 0; i < 6; ++i) {
        double Ra[9], Rn[9];
        axis_to_R(axes[i], Ra);
        mat_mul(Rprev_T, Ra, Rn);  // undo the previous axis alignment, then align z with this axis
        double p[3] = {0, 0, i == 0 ? 1.0 : 0.0};
        par = m.add_joint(par, i < 3 ? 1 : 0, names[i], Rn, p, opts);
        m.q_lo[par] = i < 3 ? -1.0 : -0.6; m.q_hi[par] = i < 3 ? 1.0 : 0.6;
        m.v_max[par] = 1.0; m.tau_max[par] = 200.0;
        m.add_frame(names[i], par, kEye, zero);
        transpose(Ra, Rprev_T);
    }
    int pelvis = par;
    {   // pelvis body on the last root joint (in the un-aligned pelvis frame)
        double c[3] = {0, 0, 0.05}, Ic[9] = {0.12, 0, 0, 0, 0.1, 0, 0, 0, 0.08};
        m.add_body(pelvis, 12.0, c, Ic, Rprev_T, zero);
        m.add_frame("pelvis", pelvis, Rprev_T, zero);
    }
    int torso = add_random_link(m, r, pelvis, "torso", opts, 0.25, 3.0);
    int remaining = ndof - 7;
    const int caps[6] = {7, 7, 4, 4, 4, 4};
    int len[6] = {0, 0, 0, 0, 0, 0};
    for (int l = 0; remaining > 0; l = (l + 1) % 6) {
        bool all_full = true;
        for (int k = 0; k < 6; ++k) all_full = all_full && len[k] >= caps[k];
        if (len[l] < caps[l] || all_full) { ++len[l]; --remaining; }
    }
    const char *limb[6] = {"larm", "rarm", "leg1", "leg2", "leg3", "leg4"};
    for (int l = 0; l < 6; ++l) {
        int p = l < 2 ? torso : pelvis;
        for (int k = 0; k < len[l]; ++k)
            p = add_random_link(m, r, p, std::string(limb[l]) + "_" + std::to_string(k + 1), opts, 0.3, l < 2 ? 1.0 : 1.5);
        if (len[l] > 0) {
            double pe[3] = {0, 0, 0.1};
            m.add_frame(std::string(limb[l]) + "_ee", p, kEye, pe);
        }
    }
    if (m.n != ndof) { err = "internal: synthetic tree size mismatch"; return MPCF_EINVAL; }
    return MPCF_OK;
}

}  // namespace mpcf
