// Loads hot_b200_plugin.so the way the reference does (SharedLibrary.cpp:5-9 dlopen RTLD_GLOBAL | RTLD_NOW; PluginManager.cpp:7-27):
// fetch `exports`, assert the API version, instantiate, registerFactories, then look the factory up.  Prints what it found;
// argv[2] = "create" additionally creates and destroys a simulation handle through the backend (needs a GPU).
#include "hot_b200_plugin.h"
#include <cstdio>
#include <cstring>
#include <dlfcn.h>

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    void* lib = dlopen(argv[1], RTLD_GLOBAL | RTLD_NOW);
    if (!lib) { std::fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
    auto* info = reinterpret_cast<ZIRAN::PluginDetails*>(dlsym(lib, "exports"));
    if (!info) { std::fprintf(stderr, "no symbol `exports`\n"); return 1; }
    std::printf("apiVersion %d\nclassName %s\npluginVersion %s\n", info->apiVersion, info->className, info->pluginVersion);
    if (info->apiVersion != ZIRAN_PLUGIN_API_VERSION) return 1;
    ZIRAN::PluginManager pm;
    pm.adopt(info);
    auto all = pm.getAll<hot_b200::Backend>();
    std::printf("plugins %d\nfactories %zu\n", pm.numPlugins(), all.size());
    if (all.size() != 1 || !all[0]) return 1;
    std::printf("supported(double,3) %d\nsupported(float,3) %d\nsupported(double,2) %d\n", (int)all[0]->supported("multigrid", true, 3),
        (int)all[0]->supported("multigrid", false, 3), (int)all[0]->supported("multigrid", true, 2));
    auto* f = dynamic_cast<ZIRAN::AFactory<hot_b200::Backend>*>(all[0]);
    if (!f) return 1;
    std::unique_ptr<hot_b200::Backend> b = f->create();
    std::printf("backend %s\nabi %s\n", b->name(), b->abiHeader());
    if (argc > 2 && !std::strcmp(argv[2], "create")) {
        hot_sim* h = b->createSimulation(1.0 / 64, 1.0, 0.6, 0);
        std::printf("handle %s\n", h ? "ok" : "null");
        if (!h) return 1;
        b->destroySimulation(h);
    }
    return 0;
}
