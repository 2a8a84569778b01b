// TEST INFRASTRUCTURE ONLY - stands in for Lib/Ziran/Math/Geometry/PartioIO.h (Partio is not in this image).  LBFGS.h only uses Partio
// in debugging output helpers that the pinned code path (LBFGS::solve) never calls; these declarations let those helpers parse.
#pragma once
#include <string>
namespace Partio {
enum ParticleAttributeType { NONE = 0, VECTOR = 1, FLOAT = 2, INT = 3 };
struct ParticleAttribute { int attributeIndex = 0; };
struct ParticlesDataMutable {
    ParticleAttribute addAttribute(const char*, ParticleAttributeType, int) { return ParticleAttribute(); }
    int addParticle() { return 0; }
    template <class T> T* dataWrite(const ParticleAttribute&, int) { return nullptr; }
    void release() {}
};
inline ParticlesDataMutable* create() { return nullptr; }
inline void write(const char*, const ParticlesDataMutable&) {}
} // namespace Partio
