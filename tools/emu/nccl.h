// tools/emu/nccl.h - type stubs for the emulated (host) build; the router's NCCL path is never run there.
#pragma once
#include <stddef.h>
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclUint64 = 5 } ncclDataType_t;
