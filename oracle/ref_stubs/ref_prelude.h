/* oracle/ref_stubs/ref_prelude.h -- TEST INFRASTRUCTURE, force-included (-include) when oracle/Makefile compiles
 * OR/OptixModel.cpp where it lies.  The reference orders tinyobj::index_t with an operator< it adds to namespace std
 * (OR/OptixModel.cpp:10-33), which only MSVC's one-phase lookup finds from inside std::less.  g++ looks the operator
 * up by ADL in namespace tinyobj, so the same lexicographic order (vertex, normal, texcoord index) is declared there;
 * the definition is in oracle/ref_scene_dump.cpp. */
#pragma once
#include "3rdParty/tiny_obj_loader.h"
namespace tinyobj {
bool operator<(const index_t& a, const index_t& b);
}
