// TEST INFRASTRUCTURE ONLY - shadows Lib/Ziran/CS/Util/ErrorContext.h: the error-context bookkeeping is a no-op.
#pragma once
#define ZIRAN_CONTEXT(...) do { } while (0)
