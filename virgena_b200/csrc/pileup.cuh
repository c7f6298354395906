#pragma once
#include "ctx.cuh"
