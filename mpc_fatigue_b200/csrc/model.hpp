// model.hpp — host-side rigid-body model: URDF reader + synthetic generators.
//
// Replaces urdf::parseURDF + pinocchio::urdf::buildModel(urdf, model, verbose) as used by the
// reference bridge (src/casadi_pinocchio_bridge.hpp:60-63): fixed base, 1-DOF joints, fixed joints
// merged into the supporting moving body, one frame per link and per joint.
#pragma once

#include <string>
#include <vector>

#include "../../include/mpcf.h"

namespace mpcf {

struct HostModel {
    int n = 0;
    std::vector<int> parent, jtype;          // [n]
    std::vector<int> jcontinuous;            // [n] 1 = URDF `continuous` joint (Pinocchio: nq = 2 (cos, sin); here: plain angle)
    std::vector<std::string> joint_names;    // [n]
    std::vector<double> Rp, pp;              // [n][9], [n][3]   joint placement in the parent joint frame
    std::vector<double> mass, mc, Io;        // [n], [n][3], [n][6]  (Io about the joint origin)
    std::vector<double> arm, fat;            // [n], [n][4]
    std::vector<double> q_lo, q_hi, v_max, tau_max;
    double grav[3] = {0.0, 0.0, -9.81};
    std::vector<std::string> frame_names;
    std::vector<int> fparent;                // [nframes], -1 = world
    std::vector<double> fR, fp;              // [nframes][9], [nframes][3]

    int add_joint(int parent_joint, int type, const std::string &name, const double R[9], const double p[3],
                  const mpcf_opts &opts);
    void add_frame(const std::string &name, int joint, const double R[9], const double p[3]);
    // add a rigid body (mass m, COM c, inertia Ic at the COM) given in a frame placed at (R, p) in joint j
    void add_body(int joint, double m, const double c[3], const double Ic[9], const double R[9], const double p[3]);
};

// returns MPCF_OK or a negative code; err receives a message
int parse_urdf(const char *xml, size_t len, const mpcf_opts &opts, HostModel &out, std::string &err);
int make_synthetic(int kind, int ndof, unsigned long long seed, const mpcf_opts &opts, HostModel &out,
                   std::string &err);

}  // namespace mpcf
