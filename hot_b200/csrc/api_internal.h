// shared by the translation units that implement the extern "C" entry points (api.cu, api_solver.cu)
#pragma once
#include "../../include/hot_b200.h"
#include "sim.h"

struct hot_sim : public hot::Sim {
    hot::DevBuf<double> stage; // AoS staging for host<->device particle marshalling
    hot::DevBuf<unsigned long long> stage_u;
    hot::DevBuf<int> stage_i;
};
