// TEST INFRASTRUCTURE ONLY - shadows Lib/Ziran/CS/Util/BinaryIO.h: the constitutive-model headers only name these in their
// (unused here) serialisation members.
#pragma once
#include <iostream>
namespace ZIRAN {
template <class T> void writeEntry(std::ostream& out, const T& x) { out.write(reinterpret_cast<const char*>(&x), sizeof(T)); }
template <class T> T readEntry(std::istream& in) { T x; in.read(reinterpret_cast<char*>(&x), sizeof(T)); return x; }
template <class T> struct RW;
template <class T> struct NoWriteTag {};
template <class T> struct CustomTypeTag {};
} // namespace ZIRAN
