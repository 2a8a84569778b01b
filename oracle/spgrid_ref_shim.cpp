// TEST INFRASTRUCTURE ONLY — not part of the product path.
//
// Thin extern "C" shim over the *reference's own* SPGrid core, compiled from the sources where they
// lie under /root/reference/Lib/SPGrid/Core (never copied into this repo).  Built by oracle/Makefile
// into oracle/_ref/libspgrid_ref.so and used by tests/ to pin the oracle's restated SPGrid addressing
// (SPGrid_Mask.h:150-166,176-189,237-245; SPGrid_Page_Map.h:61-70,90-96) bit-for-bit, and by
// tests/golden/make_spgrid_golden.py to generate the committed known-answer vectors.
#include <SPGrid/Core/SPGrid_Allocator.h>
#include <SPGrid/Core/SPGrid_Page_Map.h>
#include <array>
#include <cstring>

using namespace SPGrid;

namespace {
// Same record sizes as ZIRAN::GridState<double,3> (128 B) and GridState<float,3> (64 B)
// (Lib/MPM/MpmGrid.h:14-34); only sizeof matters for the mask.
struct Node128 { char b[128]; };
struct Node64 { char b[64]; };
using Alloc64 = SPGrid_Allocator<Node128, 3, 12>;
using Alloc32 = SPGrid_Allocator<Node64, 3, 12>;
using Mask64 = Alloc64::Array_type<>::MASK;
using Mask32 = Alloc32::Array_type<>::MASK;
}

extern "C" {

void spgrid_ref_info(int fp32, int* out /*data_bits, block_bits, xbits, ybits, zbits, elements_per_block*/)
{
    if (fp32) {
        out[0] = Mask32::data_bits; out[1] = Mask32::block_bits; out[2] = Mask32::block_xbits;
        out[3] = Mask32::block_ybits; out[4] = Mask32::block_zbits; out[5] = Mask32::elements_per_block;
    } else {
        out[0] = Mask64::data_bits; out[1] = Mask64::block_bits; out[2] = Mask64::block_xbits;
        out[3] = Mask64::block_ybits; out[4] = Mask64::block_zbits; out[5] = Mask64::elements_per_block;
    }
}

void spgrid_ref_linear_offset(int fp32, long n, const int* ijk, unsigned long long* out)
{
    for (long a = 0; a < n; ++a)
        out[a] = fp32 ? Mask32::Linear_Offset(ijk[3 * a], ijk[3 * a + 1], ijk[3 * a + 2])
                      : Mask64::Linear_Offset(ijk[3 * a], ijk[3 * a + 1], ijk[3 * a + 2]);
}

void spgrid_ref_linear_to_coord(int fp32, long n, const unsigned long long* off, int* ijk)
{
    for (long a = 0; a < n; ++a) {
        std::array<int, 3> c = fp32 ? Mask32::LinearToCoord(off[a]) : Mask64::LinearToCoord(off[a]);
        ijk[3 * a] = c[0]; ijk[3 * a + 1] = c[1]; ijk[3 * a + 2] = c[2];
    }
}

void spgrid_ref_packed_add(int fp32, long n, const unsigned long long* a, const unsigned long long* b, unsigned long long* out)
{
    for (long q = 0; q < n; ++q)
        out[q] = fp32 ? Mask32::Packed_Add(a[q], b[q]) : Mask64::Packed_Add(a[q], b[q]);
}

// Replays the page activation of MpmSimulationBase.cpp:1100-1125 through the real SPGrid_Page_Map:
// for each offset in `group_offsets` (one per particle group, in sorted order) Set_Page(offset) and
// then Set_Page of the (0/1)^3 neighbour blocks.  Returns the block list in Get_Blocks() order.
long spgrid_ref_activate(int fp32, long n_groups, const unsigned long long* group_offsets, unsigned long long* out_blocks, long cap)
{
    long count = 0;
    if (fp32) {
        static Alloc32 alloc(4096, 4096, 4096);
        SPGrid_Page_Map<12> pm(alloc);
        int x = 1 << Mask32::block_xbits, y = 1 << Mask32::block_ybits, z = 1 << Mask32::block_zbits;
        for (long g = 0; g < n_groups; ++g) {
            pm.Set_Page(group_offsets[g]);
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k)
                pm.Set_Page(Mask32::Packed_Add(group_offsets[g], Mask32::Linear_Offset(x * i, y * j, z * k)));
        }
        pm.Update_Block_Offsets();
        auto blocks = pm.Get_Blocks();
        count = blocks.second;
        for (long b = 0; b < count && b < cap; ++b) out_blocks[b] = blocks.first[b];
    } else {
        static Alloc64 alloc(4096, 4096, 4096);
        SPGrid_Page_Map<12> pm(alloc);
        int x = 1 << Mask64::block_xbits, y = 1 << Mask64::block_ybits, z = 1 << Mask64::block_zbits;
        for (long g = 0; g < n_groups; ++g) {
            pm.Set_Page(group_offsets[g]);
            for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k)
                pm.Set_Page(Mask64::Packed_Add(group_offsets[g], Mask64::Linear_Offset(x * i, y * j, z * k)));
        }
        pm.Update_Block_Offsets();
        auto blocks = pm.Get_Blocks();
        count = blocks.second;
        for (long b = 0; b < count && b < cap; ++b) out_blocks[b] = blocks.first[b];
    }
    return count;
}
}
