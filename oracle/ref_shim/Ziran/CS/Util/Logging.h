// TEST INFRASTRUCTURE ONLY - shadows Lib/Ziran/CS/Util/Logging.h: logging macros are no-ops (the rest come from the Debug.h stand-in).
#pragma once
#include <Ziran/CS/Util/Debug.h>
#ifndef ZIRAN_VERB_IF
#define ZIRAN_VERB_IF(...) do { } while (0)
#endif
