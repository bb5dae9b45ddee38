// Loaders for the reference's configuration schema (cpp/Params/*.json, cpp/src/Params/params.cpp)
// and its neural-network weight files (cpp/NNmodel/{self,env}/parameter/*.txt,
// SelfCollisionModel.cpp:60-73), plus this repo's packed fp64 weight format.
#pragma once
#include "../mpcc_types.h"
#include <map>
#include <string>
#include <vector>

namespace mpcc {

// types.h:127-150 of the reference
struct PathToJson {
    std::string param_path, cost_path, bounds_path, track_path, normalization_path, sqp_path;
};
struct ParamValue {
    std::map<std::string, double> param, cost, bounds, track, normalization, sqp;
};

// Flatten the six JSON files (with per-key overrides) into one Params record.
// Semantics follow the reference constructors, including: the solver interface's own r_ddq always
// comes from the cost FILE (osqp_interface.cpp:28,57) -> Params::r_ddq_solver.
Params load_params(const PathToJson& path, const ParamValue& overrides = ParamValue());

// Read config.json {Ts, n_sim, *_path}; paths are resolved relative to base_dir.
struct Config { double Ts; int n_sim; PathToJson paths; };
Config load_config(const std::string& config_json, const std::string& base_dir);

// A fully-connected ReLU network with the reference's [x, sin x, cos x] input encoding.
struct MlpWeights {
    std::vector<int> out_dim, in_dim;         // per layer
    std::vector<std::vector<double>> W, b;    // W[l]: out x in, row per output neuron
};
// packed format written by tools/pack_reference_assets.py
MlpWeights load_mlp_packed(const std::string& file);
// the reference's text layout: dir/weight_k.txt, dir/bias_k.txt; dims give (out,in) per layer
MlpWeights load_mlp_text(const std::string& dir, const std::vector<std::pair<int, int>>& dims);

}  // namespace mpcc
