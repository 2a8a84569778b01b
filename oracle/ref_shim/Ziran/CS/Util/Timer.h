// TEST INFRASTRUCTURE ONLY - shadows Lib/Ziran/CS/Util/Timer.h (it pulls in the logging subsystem): the timer macros are no-ops.
#pragma once
#define ZIRAN_TIMER() do { } while (0)
#define ZIRAN_QUIET_TIMER() do { } while (0)
