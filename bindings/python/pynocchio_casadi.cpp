// pynocchio_casadi.cpp — the pybind11 module the reference scripts import as `mpc_fatigue.pynocchio_casadi`
// (reference: bindings/python/pynocchio_casadi.cpp:11-16 exports generate_inv_dyn / generate_forward_kin /
// generate_jacobian traced from Pinocchio with casadi::SX, src/casadi_pinocchio_bridge.hpp:57,87,119).
//
// This body links libmpcf.so (include/mpcf.h) instead of pinocchio / casadi / urdfdom:
//   * the three generators keep their names and signatures (str urdf [, str body_name]) -> str.  Each one builds the model
//     through the C-ABI first, so a malformed URDF or an unknown frame raises here (ValueError / IndexError) instead of the
//     reference's unchecked null pointer (bridge.hpp:60-63) / oMf.at() throw (:103-107); the returned string is the token
//     that `Function.deserialize` of this module turns into the GPU-evaluating callable;
//   * `Model` is the thin batch entry point: raw device pointers (tensor.data_ptr()) + a CUDA stream in, error codes mapped
//     to Python exceptions, GIL released around every launch.
// Nothing here computes: every method is one call of the C-ABI.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "mpcf.h"

namespace py = pybind11;

namespace {

[[noreturn]] void raise(int rc)
{
    const std::string msg = mpcf_last_error();
    switch (rc) {
    case MPCF_EFRAME: throw py::index_error(msg);
    case MPCF_EINVAL: case MPCF_EPARSE: case MPCF_EJOINT: case MPCF_ELIMIT: throw py::value_error(msg);
    case MPCF_ESINGULAR: PyErr_SetString(PyExc_ZeroDivisionError, msg.c_str()); throw py::error_already_set();
    default: throw std::runtime_error(msg);
    }
}
inline void check(int rc) { if (rc < 0) raise(rc); }

using ptr = std::uintptr_t;
inline const double *cd(ptr p) { return reinterpret_cast<const double *>(p); }
inline double *md(ptr p) { return reinterpret_cast<double *>(p); }
inline void *vs(ptr p) { return reinterpret_cast<void *>(p); }

struct Model {  // RAII around the C handle
    mpcf_model *h = nullptr;
    Model(const std::string &urdf, double armature)
    {
        mpcf_opts o;
        mpcf_opts_default(&o);
        o.armature = armature;
        check(mpcf_model_create_from_urdf(urdf.data(), urdf.size(), &o, &h));
    }
    Model(const Model &) = delete;
    Model &operator=(const Model &) = delete;
    ~Model() { mpcf_model_destroy(h); }
    int nv() const { int n = 0; check(mpcf_model_info(h, nullptr, &n, nullptr, nullptr)); return n; }
    int frame_id(const std::string &name) const { int f = mpcf_frame_id(h, name.c_str()); if (f < 0) raise(f); return f; }
};

std::string json_escape(const std::string &s)
{
    static const char *hex = "0123456789abcdef";
    std::string o;
    o.reserve(s.size() + 16);
    for (unsigned char c : s) {
        switch (c) {
        case '"': o += "\\\""; break;
        case '\\': o += "\\\\"; break;
        case '\n': o += "\\n"; break;
        case '\r': o += "\\r"; break;
        case '\t': o += "\\t"; break;
        default:
            if (c < 0x20) { o += "\\u00"; o += hex[c >> 4]; o += hex[c & 15]; }
            else o += (char)c;
        }
    }
    return o;
}

// token understood by Function.deserialize (mpc_fatigue_b200/pynocchio_casadi.py)
std::string token(const char *kind, const std::string &urdf, const std::string *frame)
{
    Model probe(urdf, 0.0);  // parse now: errors surface at generation time, as in the reference
    if (frame) probe.frame_id(*frame);
    std::string s = "mpcf-function/1:{\"kind\": \"";
    s += kind;
    s += "\", \"urdf\": \"" + json_escape(urdf) + "\"";
    if (frame) s += ", \"frame\": \"" + json_escape(*frame) + "\"";
    return s + "}";
}

}  // namespace

PYBIND11_MODULE(pynocchio_casadi, m)
{
    m.doc() = "B200-native drop-in for mpc_fatigue.pynocchio_casadi (libmpcf.so behind the reference's three generators)";
    // ---- the reference's API (bindings/python/pynocchio_casadi.cpp:14-16) ----
    m.def("generate_inv_dyn", [](const std::string &urdf) { return token("inverse_dynamics", urdf, nullptr); }, py::arg("urdf_string"));
    m.def("generate_forward_kin", [](const std::string &urdf, const std::string &body) { return token("forward_kinematics", urdf, &body); },
          py::arg("urdf_string"), py::arg("body_name"));
    m.def("generate_jacobian", [](const std::string &urdf, const std::string &body) { return token("jacobian", urdf, &body); },
          py::arg("urdf_string"), py::arg("body_name"));

    // ---- thin batch entry point over the C-ABI ----
    py::class_<Model>(m, "Model")
        .def(py::init<const std::string &, double>(), py::arg("urdf"), py::arg("armature") = 0.0)
        .def_property_readonly("nv", &Model::nv)
        .def("frame_id", &Model::frame_id)
        .def("kernel_family", [](const Model &s) { return std::string(mpcf_model_kernel_family(s.h)); })
        .def("joint_names", [](const Model &s) {
            std::vector<std::string> v;
            for (int i = 0; i < s.nv(); ++i) v.emplace_back(mpcf_joint_name(s.h, i));
            return v; })
        // device pointers arrive as integers (tensor.data_ptr()), the stream as torch.cuda.current_stream().cuda_stream
        .def("rnea", [](Model &s, long U, ptr q, ptr qd, ptr qdd, ptr tau, ptr stream) {
            py::gil_scoped_release nogil;
            const int rc = mpcf_rnea_batch(s.h, U, cd(q), cd(qd), cd(qdd), md(tau), vs(stream));
            py::gil_scoped_acquire gil;
            check(rc); }, py::arg("U"), py::arg("q"), py::arg("qd"), py::arg("qdd"), py::arg("tau"), py::arg("stream") = 0)
        .def("fk", [](Model &s, int frame, long U, ptr q, ptr pos, ptr rot, ptr stream) {
            py::gil_scoped_release nogil;
            const int rc = mpcf_fk_batch(s.h, frame, U, cd(q), md(pos), md(rot), vs(stream));
            py::gil_scoped_acquire gil;
            check(rc); }, py::arg("frame"), py::arg("U"), py::arg("q"), py::arg("pos"), py::arg("rot"), py::arg("stream") = 0)
        .def("jacobian", [](Model &s, int frame, long U, ptr q, ptr J, ptr stream) {
            py::gil_scoped_release nogil;
            const int rc = mpcf_frame_jac_batch(s.h, frame, U, cd(q), md(J), vs(stream));
            py::gil_scoped_acquire gil;
            check(rc); }, py::arg("frame"), py::arg("U"), py::arg("q"), py::arg("J"), py::arg("stream") = 0)
        .def("node_eval_ref", [](Model &s, std::vector<int> frames, double wsign, long U, ptr q, ptr qd, ptr qdd, ptr W, ptr T, double h,
                                 ptr tau, ptr qnext, ptr Tnext, ptr stream) {
            py::gil_scoped_release nogil;
            const int rc = mpcf_node_eval_ref_batch(s.h, (int)frames.size(), frames.data(), wsign, U, cd(q), cd(qd), cd(qdd), cd(W), cd(T), h,
                                                    md(tau), md(qnext), md(Tnext), vs(stream));
            py::gil_scoped_acquire gil;
            check(rc); })
        .def("step_rk4", [](Model &s, long U, ptr q, ptr qd, ptr tau, ptr f, double dt, ptr qn, ptr qdn, ptr fn, ptr stream) {
            py::gil_scoped_release nogil;
            const int rc = mpcf_step_rk4_batch(s.h, U, cd(q), cd(qd), cd(tau), cd(f), dt, nullptr, md(qn), md(qdn), md(fn), vs(stream));
            py::gil_scoped_acquire gil;
            check(rc); })
        .def("step_rk4_jvp", [](Model &s, long U, ptr q, ptr qd, ptr tau, ptr f, double dt, ptr qn, ptr qdn, ptr fn, ptr jac, ptr stream) {
            py::gil_scoped_release nogil;
            const int rc = mpcf_step_rk4_jvp_batch(s.h, U, cd(q), cd(qd), cd(tau), cd(f), dt, nullptr, md(qn), md(qdn), md(fn), md(jac), vs(stream));
            py::gil_scoped_acquire gil;
            check(rc); });

    // ---- the CasADi-Function look-alike and the north-star generator live in Python (mpc_fatigue_b200/pynocchio_casadi.py) ----
    py::module_ lookalike = py::module_::import("mpc_fatigue_b200.pynocchio_casadi");
    m.attr("Function") = lookalike.attr("Function");
    m.attr("generate_fwd_dyn_fatigue_step") = lookalike.attr("generate_fwd_dyn_fatigue_step");
}
