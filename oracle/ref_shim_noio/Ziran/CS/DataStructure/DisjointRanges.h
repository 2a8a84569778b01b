// TEST INFRASTRUCTURE ONLY - stands in for Lib/Ziran/CS/DataStructure/DisjointRanges.h: only named by PlasticityApplier::applyPlasticity (see the DataManager.h stand-in)
#pragma once
namespace ZIRAN {
struct DisjointRanges {
    DisjointRanges() {}
    DisjointRanges(const DisjointRanges&, const DisjointRanges&) {}
};
} // namespace ZIRAN
