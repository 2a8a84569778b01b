// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Row (b), the plugin boundary: the REFERENCE'S OWN plugin loader - Lib/Ziran/CS/Util/{PluginManager.cpp, SharedLibrary.cpp, PluginManager.h, Plugin.h,
// Factory.h}, compiled where they lie - loads every *.so of the directory given as argv[1] exactly as main.cpp:87-93 does (PluginManager::loadAllPlugins ->
// loadPlugin: dlopen RTLD_GLOBAL | RTLD_NOW, symbol `exports`, API version assertion, initializeFunc()->registerFactories(manager)), then looks the backend
// factory up through the reference's getAll<Interface>().  include/hot_b200_plugin.h comes AFTER the reference's headers: under the same include guards it
// then adds only hot_b200::Backend, so every ZIRAN type here is the reference's definition while the plugin was compiled against the mirror.
// Built by oracle/Makefile into oracle/_ref/plugin_ref_loader; tests/test_plugin.py runs it on a directory holding hot_b200_plugin.so.
#include <cstdio>
#include <cstring>
#include <string>
#include <Ziran/CS/Util/PluginManager.h>
#include "../include/hot_b200_plugin.h"

int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    try {
        ZIRAN::PluginManager pm;
        pm.loadAllPlugins(argv[1]);
        std::printf("plugins %d\n", pm.numPlugins());
        if (pm.numPlugins() != 1) return 1;
        const ZIRAN::PluginDetails& info = pm.getPluginDetails(0);
        std::printf("apiVersion %d\nclassName %s\npluginVersion %s\n", info.apiVersion, info.className, info.pluginVersion);
        auto range = pm.getAll<hot_b200::Backend>();
        int factories = 0;
        for (auto it = range.begin(); it != range.end(); ++it) {
            ++factories;
            ZIRAN::IFactory<hot_b200::Backend>& f = *it;
            std::printf("supported(double,3) %d\nsupported(float,3) %d\nsupported(double,2) %d\n", (int)f.supported("multigrid", true, 3),
                (int)f.supported("multigrid", false, 3), (int)f.supported("multigrid", true, 2));
            ZIRAN::AFactory<hot_b200::Backend>* af = it;
            if (!af) return 1;
            std::unique_ptr<hot_b200::Backend> b = af->create();
            std::printf("backend %s\nabi %s\n", b->name(), b->abiHeader());
        }
        std::printf("factories %d\n", factories);
        return factories == 1 ? 0 : 1;
    }
    catch (std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
}
