/* oracle/shim/pfc/pfc.h -- TEST INFRASTRUCTURE. The reference's util.h includes "../../pfc/pfc.h" (the foobar2000
 * SDK's utility library, not part of the reference tree) and uses exactly one thing from it: pfc::max_t. */
#pragma once
namespace pfc {
template <typename T> inline T max_t(const T &a, const T &b) { return a > b ? a : b; }
template <typename T> inline T min_t(const T &a, const T &b) { return a < b ? a : b; }
}
