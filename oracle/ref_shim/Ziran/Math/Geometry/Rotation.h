// TEST INFRASTRUCTURE ONLY - stands in for Lib/Ziran/Math/Geometry/Rotation.h, which Lib/Ziran/Physics/PlasticityApplier.h includes without using anything of it
#pragma once
