// forwards to the single-file re-declaration of the Renderer boundary
#pragma once
#include "Render/RenderAPI.h"
