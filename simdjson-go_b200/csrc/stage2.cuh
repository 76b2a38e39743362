#pragma once
#include "common.cuh"
