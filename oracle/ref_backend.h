// ORACLE (test infrastructure, not product code).
// Reference-side backend of the trace player: pulls in the UNMODIFIED reference headers from
// /root/reference/src so that the player drives the reference's own RendererSoft.  Nothing here is
// copied from the reference; it only includes it.  Built by oracle/Makefile into oracle/_ref/.
#pragma once
#include "Render/Renderer.h"
#include "Render/Software/RendererSoft.h"
#include "Render/Software/TextureSoft.h"
