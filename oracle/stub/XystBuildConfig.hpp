// Test infrastructure only: stands in for the cmake-generated build config of
// the reference (cmake/ConfigureDataLayout.cmake default: field-major layout).
#pragma once
#define FIELD_DATA_LAYOUT_AS_FIELD_MAJOR
