// Test scaffolding only: stands in for googletest's gtest_prod.h, which the
// reference includes (include/ilqr.h:12) but fetches from the network at build
// time (CMakeLists.txt:67-76). Makes the FRIEND_TEST hooks (ilqr.h:103-106)
// name classes that oracle/ref_harness.cpp defines.
#pragma once
#define FRIEND_TEST(test_case_name, test_name) friend class test_case_name##_##test_name##_Test
