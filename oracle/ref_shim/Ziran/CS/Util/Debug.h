// TEST INFRASTRUCTURE ONLY - shadows Lib/Ziran/CS/Util/Debug.h (which pulls in the logging subsystem): ZIRAN_ASSERT throws like the
// reference's (Debug.h:19-42), logging macros are no-ops.
#pragma once
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#define ZIRAN_ASSERT(cond, ...) do { if (!(cond)) throw std::runtime_error("Assertion failed"); } while (0)
#define ZIRAN_INFO(...) do { } while (0)
#define ZIRAN_WARN(...) do { } while (0)
#define ZIRAN_DEBUG(...) do { } while (0)
#define ZIRAN_VERB(...) do { } while (0)
#define ZIRAN_ERR(...) do { } while (0)
