// TEST INFRASTRUCTURE ONLY - shadows Lib/Ziran/CS/Util/Timer.h (it pulls in the logging subsystem): the timer macros are no-ops and
// ZIRAN::Timer (LBFGS.h times its low-rank updates with it) measures nothing.
#pragma once
#define ZIRAN_TIMER() do { } while (0)
#define ZIRAN_QUIET_TIMER() do { } while (0)
namespace ZIRAN {
struct Timer {
    struct Duration { double count() const { return 0.0; } };
    void start() {}
    Duration click(bool = false) { return Duration(); }
};
} // namespace ZIRAN
