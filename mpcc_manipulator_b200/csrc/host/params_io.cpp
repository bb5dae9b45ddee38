#include "params_io.h"
#include "json_min.h"
#include <cstring>
#include <fstream>
#include <stdexcept>

namespace mpcc {

namespace {
double pick(const json::Value& j, const std::map<std::string, double>& ov, const std::string& key) {
    auto it = ov.find(key);
    return it == ov.end() ? j.number(key) : it->second;
}
const char* X_NAMES[NX] = {"q1", "q2", "q3", "q4", "q5", "q6", "q7", "s", "vs"};
const char* U_NAMES[NU] = {"dq1", "dq2", "dq3", "dq4", "dq5", "dq6", "dq7", "dVs"};
}  // namespace

Params load_params(const PathToJson& path, const ParamValue& ov) {
    Params p;
    std::memset(&p, 0, sizeof(p));
    {  // Param (params.cpp:24-85)
        json::Value j = json::parse_file(path.param_path);
        p.max_dist_proj = pick(j, ov.param, "max_dist_proj");
        p.desired_ee_velocity = pick(j, ov.param, "desired_ee_velocity");
        p.s_trust_region = pick(j, ov.param, "s_trust_region");
        p.deacc_ratio = pick(j, ov.param, "deaccelerate_ratio");
        p.tol_sing = pick(j, ov.param, "tol_sing");
        p.tol_selcol = pick(j, ov.param, "tol_selcol");
        p.tol_envcol = pick(j, ov.param, "tol_envcol");
    }
    {  // CostParam (params.cpp:87-178)
        json::Value j = json::parse_file(path.cost_path);
        p.q_c = pick(j, ov.cost, "qC");
        p.q_c_N_mult = pick(j, ov.cost, "qCNmult");
        p.q_l = pick(j, ov.cost, "qL");
        p.q_vs = pick(j, ov.cost, "qVs");
        p.q_ori = pick(j, ov.cost, "qOri");
        p.q_sing = pick(j, ov.cost, "qSing");
        p.r_dq = pick(j, ov.cost, "rdq");
        p.r_ddq = pick(j, ov.cost, "rddq");
        p.r_dVs = pick(j, ov.cost, "rdVs");
        p.q_c_red_ratio = pick(j, ov.cost, "qC_reduction_ratio");
        p.q_l_inc_ratio = pick(j, ov.cost, "qL_increase_ratio");
        p.q_ori_red_ratio = pick(j, ov.cost, "qOri_reduction_ratio");
        p.r_ddq_solver = j.number("rddq");  // file value only
    }
    {  // BoundsParam (params.cpp:180-306).  The reference's Bounds object is always built from the
       // file (osqp_interface.cpp:25,54,99), so overrides in ParamValue::bounds are ignored there too.
        json::Value j = json::parse_file(path.bounds_path);
        for (int i = 0; i < NX; i++) { p.lx[i] = j.number(std::string(X_NAMES[i]) + "l"); p.ux[i] = j.number(std::string(X_NAMES[i]) + "u"); }
        for (int i = 0; i < NU; i++) { p.lu[i] = j.number(std::string(U_NAMES[i]) + "l"); p.uu[i] = j.number(std::string(U_NAMES[i]) + "u"); }
        for (int i = 0; i < DOF; i++) {
            std::string k = "ddq" + std::to_string(i + 1);
            p.ldd[i] = j.number(k + "l");
            p.udd[i] = j.number(k + "u");
        }
    }
    {  // NormalizationParam (params.cpp:308-402)
        json::Value j = json::parse_file(path.normalization_path);
        for (int i = 0; i < NX; i++) p.Tx[i] = pick(j, ov.normalization, X_NAMES[i]);
        for (int i = 0; i < NU; i++) p.Tu[i] = pick(j, ov.normalization, U_NAMES[i]);
    }
    {  // SQPParam (params.cpp:404-448)
        json::Value j = json::parse_file(path.sqp_path);
        p.eps_prim = pick(j, ov.sqp, "eps_prim");
        p.eps_dual = pick(j, ov.sqp, "eps_dual");
        p.max_iter = (double)(int)pick(j, ov.sqp, "max_iter");
        p.line_search_max_iter = (double)(int)pick(j, ov.sqp, "line_search_max_iter");
        p.do_SOC = pick(j, ov.sqp, "do_SOC") != 0 ? 1.0 : 0.0;
        p.use_BFGS = pick(j, ov.sqp, "use_BFGS") != 0 ? 1.0 : 0.0;
        p.line_search_tau = pick(j, ov.sqp, "line_search_tau");
        p.line_search_eta = pick(j, ov.sqp, "line_search_eta");
        p.line_search_rho = pick(j, ov.sqp, "line_search_rho");
    }
    return p;
}

Config load_config(const std::string& config_json, const std::string& base_dir) {
    json::Value j = json::parse_file(config_json);
    Config c;
    c.Ts = j.number("Ts");
    c.n_sim = (int)j.number("n_sim");
    std::string b = base_dir;
    if (!b.empty() && b.back() != '/') b += '/';
    c.paths.param_path = b + j.string("model_path");
    c.paths.cost_path = b + j.string("cost_path");
    c.paths.bounds_path = b + j.string("bounds_path");
    c.paths.track_path = b + j.string("track_path");
    c.paths.normalization_path = b + j.string("normalization_path");
    c.paths.sqp_path = b + j.string("sqp_path");
    return c;
}

MlpWeights load_mlp_packed(const std::string& file) {
    std::ifstream f(file, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("nn: cannot open '" + file + "'");
    char magic[8];
    f.read(magic, 8);
    if (std::memcmp(magic, "MPCCNN1\0", 8) != 0) throw std::runtime_error("nn: bad magic in '" + file + "'");
    int32_t nl = 0;
    f.read((char*)&nl, 4);
    if (nl <= 0 || nl > 16) throw std::runtime_error("nn: bad layer count in '" + file + "'");
    MlpWeights m;
    for (int l = 0; l < nl; l++) {
        int32_t d[2];
        f.read((char*)d, 8);
        m.out_dim.push_back(d[0]);
        m.in_dim.push_back(d[1]);
    }
    for (int l = 0; l < nl; l++) {
        m.W.emplace_back((size_t)m.out_dim[l] * m.in_dim[l]);
        m.b.emplace_back((size_t)m.out_dim[l]);
        f.read((char*)m.W[l].data(), (std::streamsize)(m.W[l].size() * 8));
        f.read((char*)m.b[l].data(), (std::streamsize)(m.b[l].size() * 8));
    }
    if (!f) throw std::runtime_error("nn: truncated file '" + file + "'");
    return m;
}

MlpWeights load_mlp_text(const std::string& dir, const std::vector<std::pair<int, int>>& dims) {
    MlpWeights m;
    std::string d = dir;
    if (!d.empty() && d.back() != '/') d += '/';
    for (size_t l = 0; l < dims.size(); l++) {
        m.out_dim.push_back(dims[l].first);
        m.in_dim.push_back(dims[l].second);
        std::ifstream fw(d + "weight_" + std::to_string(l) + ".txt"), fb(d + "bias_" + std::to_string(l) + ".txt");
        if (!fw.is_open() || !fb.is_open()) throw std::runtime_error("nn: cannot open layer " + std::to_string(l) + " under '" + dir + "'");
        m.W.emplace_back((size_t)dims[l].first * dims[l].second);
        m.b.emplace_back((size_t)dims[l].first);
        for (double& v : m.W[l]) if (!(fw >> v)) throw std::runtime_error("nn: short weight file, layer " + std::to_string(l));
        for (double& v : m.b[l]) if (!(fb >> v)) throw std::runtime_error("nn: short bias file, layer " + std::to_string(l));
    }
    return m;
}

}  // namespace mpcc
