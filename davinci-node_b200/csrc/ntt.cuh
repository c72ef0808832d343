#pragma once
#include "field.cuh"
namespace b200 {}
