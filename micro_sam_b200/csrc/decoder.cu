// placeholder until the decoder lands
#include "engine.h"
namespace msam {
int Engine::finalize_decoder() { return 0; }
}
