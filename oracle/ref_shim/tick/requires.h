// TEST INFRASTRUCTURE ONLY - stand-in for Tick's TICK_REQUIRES (https://github.com/pfultz2/Tick, pinned @b82af54 upstream, absent here):
// SFINAE constraints on template parameters, nothing arithmetic.
#pragma once
#include <type_traits>
#define TICK_REQUIRES(...) bool TickPrivateBool__ = true, typename std::enable_if<(TickPrivateBool__ && (__VA_ARGS__)), int>::type = 0
#define TICK_MEMBER_REQUIRES(...) template <bool TickPrivateBool__ = true, typename std::enable_if<(TickPrivateBool__ && (__VA_ARGS__)), int>::type = 0>
