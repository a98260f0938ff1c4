// RendererCUDA-side backend of the trace player (the product harness).
#pragma once
#include "Render/Renderer.h"
#include "Render/CUDA/RendererCUDA.h"
