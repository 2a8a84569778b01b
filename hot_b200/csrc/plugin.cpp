// hot_b200_plugin.so: the shared-object plugin the reference's loader picks up from ZIRAN_PLUGIN_DIR (main.cpp:87-93 ->
// PluginManager::loadAllPlugins -> loadPlugin, PluginManager.cpp:7-27).  Exports `extern "C" ZIRAN::PluginDetails exports`
// (Plugin.h:21-43) and registers a factory for hot_b200::Backend.  Host code only: the CUDA work sits behind the C ABI of
// libhot_b200.so, which this object links.
#include "../../include/hot_b200_plugin.h"

namespace {

class BackendB200 final : public hot_b200::Backend {
public:
    const char* name() const override { return "hot_b200: implicit-MPM hot path on B200 (sm_100a)"; }
    const char* abiHeader() const override { return "hot_b200.h"; }
    hot_sim* createSimulation(double dx, double apic_rpic_ratio, double cfl, int device) override { return hot_create(dx, apic_rpic_ratio, cfl, device); }
    void destroySimulation(hot_sim* h) override { hot_destroy(h); }
};

class BackendFactory final : public ZIRAN::Factory<BackendB200, hot_b200::Backend> {
public:
    // the reference binary is T = double, dim = 3 (Projects/multigrid/main.cpp:12-13); any simulation name
    bool supported(const char*, bool use_double, int dimension) override { return use_double && dimension == 3; }
};

class HotB200Plugin final : public ZIRAN::PluginBase {
public:
    void registerFactories(ZIRAN::PluginManager& manager) override
    {
        manager.registerFactory<hot_b200::Backend>(std::unique_ptr<ZIRAN::IFactory<hot_b200::Backend>>(new BackendFactory));
    }
};

} // namespace

ZIRAN_PLUGIN(HotB200Plugin, "2")
