// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// f4, restart files: the REFERENCE'S OWN attribute store and binary writers - Lib/Ziran/CS/DataStructure/{DataManager.h, DataArray.h, DataArrayBase.h,
// DisjointRanges.h}, Lib/Ziran/CS/Util/BinaryIO.h and CorotatedIsotropic::write - compiled where they lie (this library is built WITHOUT the inert
// stand-ins of ref_shim_noio/ that shadow those headers for the other libraries).  zr_restart_write fills a DataManager with the particle arrays
// of this path ("X", "V", "m", "element measure", "F", "CorotatedIsotropic") and serialises it the way MpmSimulationBase::writeState does for an MPM
// scene (Lib/MPM/MpmSimulationBase.cpp:755-770 -> Scene::writeState, Lib/Ziran/Sim/Scene.h:189-206: particles.writeData, no element managers, the
// two empty mesh index vectors); zr_restart_read runs DataManager::readData on a byte string and hands the arrays back.
// Built by oracle/Makefile into oracle/_ref/librestart_ref.so; tests/golden/make_restart_golden.py, tests/test_restart_ref.py.
#include <functional>
#include <cstring>
#include <sstream>
#include <string>
#include <Ziran/CS/Util/Forward.h>
#include <tbb/tbb.h>
#include <Ziran/CS/DataStructure/DataManager.h>
#include <Ziran/Physics/ConstitutiveModel/HyperelasticConstitutiveModel.h>
#include <Ziran/Physics/ConstitutiveModel/CorotatedIsotropic.h>

using namespace ZIRAN;
typedef double T;
typedef Vector<T, 3> TV;
typedef Matrix<T, 3, 3> TM;
typedef CorotatedIsotropic<T, 3> Model;

extern "C" {

long zr_restart_write(long n, const double* X, const double* V, const double* m, const double* vol, const double* F, const double* mu, const double* lam,
    unsigned char* out, long cap)
{
    DataManager particles;
    Range r{0, (int)n};
    StdVector<TV> x(n), v(n);
    StdVector<T> mm(m, m + n), vv(vol, vol + n);
    StdVector<TM> f(n);
    StdVector<Model> models(n);
    for (long i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) { x[i](d) = X[3 * i + d]; v[i](d) = V[3 * i + d]; }
        for (int q = 0; q < 9; ++q) f[i](q) = F[9 * i + q];
        models[i].mu = mu[i];
        models[i].lambda = lam[i];
    }
    particles.add(AttributeName<TV>("X"), r, std::move(x));
    particles.add(AttributeName<TV>("V"), r, std::move(v));
    particles.add(AttributeName<T>("m"), r, std::move(mm));
    particles.add(AttributeName<T>("element measure"), r, std::move(vv));
    particles.add(AttributeName<TM>("F"), r, std::move(f));
    particles.add(AttributeName<Model>(Model::name()), r, std::move(models));
    std::ostringstream os(std::ios::binary);
    particles.writeData(os);
    writeSTDVector(os, StdVector<Vector<int, 3>>()); // trimesh_to_write.indices
    writeSTDVector(os, StdVector<Vector<int, 2>>()); // segmesh_to_write.indices
    const std::string s = os.str();
    if ((long)s.size() > cap) return -(long)s.size();
    std::memcpy(out, s.data(), s.size());
    return (long)s.size();
}

// DataManager::readData (DataManager.h:280-293) on `bytes` (empty arrays of the right types created first, as the reference requires); returns the
// particle count, or -1 when the reference's reader throws
long zr_restart_read(const unsigned char* bytes, long len, long cap, double* X, double* V, double* m, double* vol, double* F, double* mu, double* lam)
{
    try {
        DataManager particles;
        auto& x = particles.add(AttributeName<TV>("X"));
        auto& v = particles.add(AttributeName<TV>("V"));
        auto& mm = particles.add(AttributeName<T>("m"));
        auto& vv = particles.add(AttributeName<T>("element measure"));
        auto& f = particles.add(AttributeName<TM>("F"));
        auto& mo = particles.add(AttributeName<Model>(Model::name()));
        std::istringstream is(std::string((const char*)bytes, (size_t)len), std::ios::binary);
        particles.readData(is);
        StdVector<Vector<int, 3>> tri;
        StdVector<Vector<int, 2>> seg;
        readSTDVector(is, tri);
        readSTDVector(is, seg);
        const long n = particles.count;
        if (n > cap) return -2;
        for (long i = 0; i < n; ++i) {
            for (int d = 0; d < 3; ++d) { X[3 * i + d] = x.array[i](d); V[3 * i + d] = v.array[i](d); }
            for (int q = 0; q < 9; ++q) F[9 * i + q] = f.array[i](q);
            m[i] = mm.array[i]; vol[i] = vv.array[i];
            mu[i] = mo.array[i].mu; lam[i] = mo.array[i].lambda;
        }
        return n;
    }
    catch (...) {
        return -1;
    }
}

} // extern "C"
