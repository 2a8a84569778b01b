// shared by the translation units that implement the extern "C" entry points (api.cu, api_solver.cu)
#pragma once
#include "../../include/hot_b200.h"
#include "sim.h"

struct hot_sim : public hot::Sim {
    hot::DevBuf<double> stage; // AoS staging for host<->device particle marshalling
    hot::DevBuf<unsigned long long> stage_u;
    hot::DevBuf<int> stage_i;
    // pipelined host <-> device particle state (hot_upload_state_async ... hot_wait_download): own copy streams so that the
    // upload of step k+1 and the download of step k run on the two PCIe directions while step k computes
    hot::DevBuf<double> stage_in, stage_out;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in_done = nullptr, ev_in_free = nullptr, ev_out_ready = nullptr, ev_out_done = nullptr;
    bool in_pending = false, in_free_recorded = false, out_pending = false;
};
