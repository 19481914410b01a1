// TEST STUB - not XLA.  The adapter includes xla/ffi/api/c_api.h for completeness; nothing from it is used directly.
#pragma once
