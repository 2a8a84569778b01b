// The restart-file writer / reader of include/hot_b200_host.hpp (writeRestart / readRestart, the serialisation under MpmSimulationB200::writeState /
// readState) on host arrays, for the comparison with the reference's own DataManager / BinaryIO code (oracle/restart_ref_shim.cpp,
// tests/test_restart_ref.py).  No device call is made.
//   restart_ref write <arrays.bin> <restart.dat>     arrays.bin: int64 n, X[3n] V[3n] m[n] vol[n] F[9n] mu[n] lam[n]   (no trailing APIC block)
//   restart_ref read  <restart.dat> <arrays.bin>
#include "hot_b200_host.hpp"
#include <cstdio>
#include <fstream>

using namespace hot_b200;

int main(int argc, char** argv)
{
    if (argc < 4) return 2;
    try {
        if (std::string(argv[1]) == "write") {
            std::ifstream in(argv[2], std::ios::binary);
            long long n = 0;
            in.read(reinterpret_cast<char*>(&n), 8);
            auto rd = [&](size_t k) { std::vector<double> v(k); in.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(8 * k)); return v; };
            auto X = rd(3 * n), V = rd(3 * n), m = rd(n), vol = rd(n), F = rd(9 * n), mu = rd(n), lam = rd(n);
            if (!in) return 3;
            std::ofstream out(argv[3], std::ios::binary);
            HOTSettings::project = true; // --project, as in the HOT runs (and the default of the reference's model object, CorotatedIsotropic.h:61)
            writeRestart(out, (long)n, X.data(), V.data(), m.data(), vol.data(), F.data(), mu.data(), lam.data(), nullptr);
        }
        else {
            std::ifstream in(argv[2], std::ios::binary);
            RestartArrays a = readRestart(in);
            std::ofstream out(argv[3], std::ios::binary);
            long long n = a.n;
            out.write(reinterpret_cast<const char*>(&n), 8);
            for (const std::vector<double>* v : {&a.X, &a.V, &a.m, &a.vol, &a.F, &a.mu, &a.lam}) out.write(reinterpret_cast<const char*>(v->data()), (std::streamsize)(8 * v->size()));
        }
    }
    catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
