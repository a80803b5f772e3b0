// placeholder until the physics kernel lands (next commit): fails loudly
#include "internal.h"
namespace dyros {
int physics_configure(Sim*) { return 0; }
int launch_simulate(Sim*, int, const float*, cudaStream_t) { set_error("physics kernel not built"); return 1; }
int launch_refresh_rigid_body_state(Sim*, cudaStream_t) { set_error("physics kernel not built"); return 1; }
}
