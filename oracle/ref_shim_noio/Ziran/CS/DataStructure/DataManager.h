// TEST INFRASTRUCTURE ONLY - stands in for Lib/Ziran/CS/DataStructure/DataManager.h (the reference's attribute store: ranges, hash tables, TBB).
// Lib/Ziran/Physics/PlasticityApplier.{h,cpp} is compiled where it lies for its return mappings (SnowPlasticity::projectStrain,
// VonMisesFixedCorotated::projectStrain); the explicit instantiations at the end of that file also instantiate
// PlasticityApplier::applyPlasticity, a loop over a DataManager subset - the declarations below let it compile, it is never called
// (oracle/plasticity_ref_shim.cpp calls the return mappings per particle, which is all that loop does: PlasticityApplier.h:38-50).
#pragma once
#include <string>
#include <tuple>
#include <vector>
#include <Ziran/CS/DataStructure/DisjointRanges.h>
namespace ZIRAN {
template <class Type>
struct AttributeName {
    std::string name;
    AttributeName(const std::string& n) : name(n) {}
    AttributeName(const char* n) : name(n) {}
};
template <class... Types>
struct SubsetIterStandIn {
    std::tuple<Types*...> p;
    explicit operator bool() const { return false; }
    SubsetIterStandIn& operator++() { return *this; }
    template <int I>
    typename std::tuple_element<I, std::tuple<Types...>>::type& get() { return *std::get<I>(p); }
};
class DataManager {
public:
    // the two scalar attributes Projects/multigrid/ImplicitSolver.h reads through particles.DataManager::get(AttributeName<T>(..)) ("element measure" :500,618
    // and "m" :676), served from the arrays of the stand-in simulation of oracle/implicit_ref_shim.cpp
    std::vector<double>* measure = nullptr;
    std::vector<double>* m = nullptr;
    std::vector<double>& get(const AttributeName<double>& a) { return a.name == "m" ? *m : *measure; }
    template <class Type>
    bool exist(const AttributeName<Type>&) const { return false; }
    template <class... Types>
    DisjointRanges commonRanges(const AttributeName<Types>&...) const { return DisjointRanges(); }
    template <class... Types>
    SubsetIterStandIn<Types...> subsetIter(const DisjointRanges&, const AttributeName<Types>&...) { return SubsetIterStandIn<Types...>(); }
};
} // namespace ZIRAN
