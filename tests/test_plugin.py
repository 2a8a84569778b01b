"""The reference's plugin ABI (Lib/Ziran/CS/Util/Plugin.h:11-43, loader PluginManager.cpp:7-27): hot_b200_plugin.so is dlopen'ed by a
small C++ host that does what the reference's loader does (tests/cpp/plugin_load.cpp).  No GPU needed to load it and to look up
its factory; creating a simulation handle through the backend is the -m gpu part."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PLUGIN = os.path.join(ROOT, "hot_b200", "lib", "hot_b200_plugin.so")


@pytest.fixture(scope="module")
def loader(tmp_path_factory):
    if not os.path.exists(PLUGIN):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "hot_b200", "csrc")])
    exe = str(tmp_path_factory.mktemp("cpp") / "plugin_load")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "plugin_load.cpp"), "-o", exe, "-ldl"])
    return exe


def _fields(out):
    return dict(line.split(" ", 1) for line in out.strip().splitlines())


def test_plugin_exports_and_factory(loader):
    f = _fields(subprocess.check_output([loader, PLUGIN], text=True))
    assert f["apiVersion"] == "2" and f["className"] == "HotB200Plugin"
    assert f["plugins"] == "1" and f["factories"] == "1"
    assert f["supported(double,3)"] == "1" and f["supported(float,3)"] == "0" and f["supported(double,2)"] == "0"
    assert f["abi"] == "hot_b200.h"


def test_exports_symbol_is_a_data_object():
    out = subprocess.check_output(["nm", "-D", "--defined-only", PLUGIN], text=True)
    assert any(line.split()[-1] == "exports" and line.split()[-2] in "DdBb" for line in out.splitlines())
    assert any(line.split()[-1] == "getHotB200Plugin" for line in out.splitlines())


@pytest.mark.gpu
def test_backend_creates_a_handle(loader):
    f = _fields(subprocess.check_output([loader, PLUGIN, "create"], text=True))
    assert f["handle"] == "ok"


REF_LOADER = os.path.join(ROOT, "oracle", "_ref", "plugin_ref_loader")


@pytest.mark.skipif(not os.path.exists(REF_LOADER), reason="oracle/_ref/plugin_ref_loader not built (needs /root/reference)")
def test_reference_plugin_manager_loads_the_plugin(tmp_path):
    """the REFERENCE'S OWN loader - Lib/Ziran/CS/Util/{PluginManager.cpp, SharedLibrary.cpp} compiled where they lie (oracle/plugin_ref_loader.cpp) - runs
    loadAllPlugins on a plugin directory holding hot_b200_plugin.so (a copy, like <build>/Plugins; libhot_b200.so on the library path), asserts the API
    version, lets the plugin register its factory in the reference's PluginManager and retrieves it through the reference's getAll<Interface>()"""
    import shutil
    plugins = tmp_path / "Plugins"
    plugins.mkdir()
    shutil.copy(PLUGIN, plugins / "hot_b200_plugin.so")
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "hot_b200", "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    f = _fields(subprocess.check_output([REF_LOADER, str(plugins)], text=True, env=env))
    assert f["plugins"] == "1" and f["factories"] == "1"
    assert f["apiVersion"] == "2" and f["className"] == "HotB200Plugin"
    assert f["supported(double,3)"] == "1" and f["supported(float,3)"] == "0" and f["supported(double,2)"] == "0"
    assert f["abi"] == "hot_b200.h"
