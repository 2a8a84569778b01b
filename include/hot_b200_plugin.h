// Plugin side of hot_b200: the reference's shared-object plugin ABI (Lib/Ziran/CS/Util/Plugin.h:11-43, Factory.h:6-41,
// PluginManager.h:13-43; loader PluginManager.cpp:7-27: every *.so of ZIRAN_PLUGIN_DIR is dlopen'ed RTLD_GLOBAL | RTLD_NOW, the
// symbol `exports` is read, apiVersion asserted == 2 and initializeFunc()->registerFactories(manager) called).
//
// hot_b200/lib/hot_b200_plugin.so exports that symbol; its registerFactories registers ONE factory for the interface
// hot_b200::Backend below (a thin C++ face of the C ABI of hot_b200.h), supported() for (any simulation name, double, 3-D).
//
// The declarations in namespace ZIRAN mirror the reference's headers member for member (the ABI is their layout: vtable order
// of PluginBase / FactoryBase, the PluginDetails struct, the data members of PluginManager) under THE SAME INCLUDE GUARDS, so a
// translation unit inside the reference tree that has already included <Ziran/CS/Util/PluginManager.h> uses the reference's own
// definitions and this file adds only hot_b200::Backend.
#ifndef HOT_B200_PLUGIN_H
#define HOT_B200_PLUGIN_H
#include "hot_b200.h"
#include <cassert>
#include <memory>
#include <typeindex>
#include <typeinfo>
#include <unordered_map>
#include <vector>

#ifndef PLATFORM_SPECIFIC_H
#define PLATFORM_SPECIFIC_H
#define ZIRAN_FORCE_INLINE __attribute__((always_inline))
#define ZIRAN_EXPORT __attribute__((visibility("default")))
#define ZIRAN_LOCAL __attribute__((visibility("hidden")))
#endif

#ifndef FACTORY_H
#define FACTORY_H value
namespace ZIRAN {
class ZIRAN_EXPORT FactoryBase { // Factory.h:6-11
public:
    virtual ~FactoryBase() {}
    virtual const std::type_info& createdTypeInfo() const = 0;
    virtual bool supported(const char* sim_name, bool use_double, int dimension) = 0;
};
template <class Interface>
class ZIRAN_EXPORT IFactory : public FactoryBase { // :13-17
public:
    virtual ~IFactory() {}
};
template <class Interface, class... Args>
class ZIRAN_EXPORT AFactory : public virtual IFactory<Interface> { // :19-25
public:
    virtual ~AFactory() {}
    virtual std::unique_ptr<Interface> create(Args... args) const = 0;
};
template <class Derived, class Interface, class... Args>
class ZIRAN_EXPORT Factory : public AFactory<Interface, Args...> { // :27-41
public:
    virtual ~Factory() {}
    virtual std::unique_ptr<Interface> create(Args... args) const override { return std::make_unique<Derived>(args...); }
    const std::type_info& createdTypeInfo() const override { return typeid(Derived); }
};
} // namespace ZIRAN
#endif

#ifndef PLUGIN_H
#define PLUGIN_H
namespace ZIRAN {
class PluginManager;
#define ZIRAN_PLUGIN_API_VERSION 2
class ZIRAN_EXPORT PluginBase { // Plugin.h:13-18
public:
    PluginBase() {}
    virtual ~PluginBase() {}
    virtual void registerFactories(PluginManager& manager) = 0;
};
struct PluginDetails { // :21-27
    int apiVersion;
    const char* fileName;
    const char* className;
    const char* pluginVersion;
    PluginBase* (*initializeFunc)();
};
#define ZIRAN_PLUGIN(classType, pluginVersion)       \
    extern "C" {                                     \
    ZIRAN_EXPORT ZIRAN::PluginBase* get##classType() \
    {                                                \
        static classType singleton;                  \
        return &singleton;                           \
    }                                                \
    ZIRAN_EXPORT ZIRAN::PluginDetails exports = {    \
        ZIRAN_PLUGIN_API_VERSION,                    \
        __FILE__,                                    \
        #classType,                                  \
        pluginVersion,                               \
        get##classType                               \
    };                                               \
    }
} // namespace ZIRAN
#endif

#ifndef PLUGIN_MANAGER_H
#define PLUGIN_MANAGER_H
namespace ZIRAN {
class SharedLibrary; // SharedLibrary.h: only the vector's element type
// data members and registerFactory of PluginManager.h:13-27 (what a plugin touches); loading stays with the host application
class ZIRAN_EXPORT PluginManager {
    using TypeMap = std::unordered_multimap<std::type_index, std::unique_ptr<FactoryBase>>;
    std::vector<SharedLibrary*> shared_libraries_layout_only; // std::vector<SharedLibrary> in the reference: same three words
    std::vector<PluginBase*> plugins;
    std::vector<PluginDetails*> plugin_details;
    TypeMap data;

public:
    template <class Interface>
    void registerFactory(std::unique_ptr<IFactory<Interface>>&& factory) { data.emplace(typeid(Interface), std::move(factory)); }
    // (host side of this mirror, used by tests/cpp/plugin_load.cpp the way PluginManager.cpp:7-27 uses the reference's)
    void adopt(PluginDetails* info)
    {
        plugin_details.push_back(info);
        plugins.push_back(info->initializeFunc());
        plugins.back()->registerFactories(*this);
    }
    int numPlugins() const { return (int)plugins.size(); }
    template <class Interface>
    std::vector<IFactory<Interface>*> getAll() const
    {
        std::vector<IFactory<Interface>*> out;
        auto range = data.equal_range(typeid(Interface));
        for (auto it = range.first; it != range.second; ++it) out.push_back(dynamic_cast<IFactory<Interface>*>(it->second.get()));
        return out;
    }
};
} // namespace ZIRAN
#endif

namespace hot_b200 {
// What the plugin's factory creates: the B200 backend of the implicit-MPM hot path as one object per GPU.
class Backend {
public:
    virtual ~Backend() {}
    virtual const char* name() const = 0;
    virtual const char* abiHeader() const = 0;                    // "hot_b200.h": the C ABI the handle speaks
    // hot_create / hot_destroy; the caller drives the handle through hot_b200.h or include/hot_b200_host.hpp
    virtual hot_sim* createSimulation(double dx, double apic_rpic_ratio, double cfl, int device) = 0;
    virtual void destroySimulation(hot_sim* h) = 0;
};
} // namespace hot_b200
#endif
