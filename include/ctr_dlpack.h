/* ctr_dlpack.h -- the subset of the DLPack ABI (dmlc/dlpack, v0.8 layout) that
 * libctradon's *_dl entry points accept.  Field order and widths follow the public
 * specification so a DLTensor / DLManagedTensor produced by torch.utils.dlpack,
 * tf.experimental.dlpack or cupy can be passed as is.  If <dlpack/dlpack.h> was
 * included first, its definitions are used instead. */
#ifndef CTR_DLPACK_H_
#define CTR_DLPACK_H_
#include <stdint.h>

#ifndef DLPACK_DLPACK_H_
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3, kDLCUDAManaged = 13 } DLDeviceType;
typedef struct { DLDeviceType device_type; int32_t device_id; } DLDevice;
typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U, kDLBfloat = 4U } DLDataTypeCode;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;
typedef struct {
    void*      data;
    DLDevice   device;
    int32_t    ndim;
    DLDataType dtype;
    int64_t*   shape;
    int64_t*   strides;      /* in elements; NULL = compact row-major */
    uint64_t   byte_offset;
} DLTensor;
typedef struct DLManagedTensor {
    DLTensor dl_tensor;
    void*    manager_ctx;
    void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;
#endif /* DLPACK_DLPACK_H_ */
#endif /* CTR_DLPACK_H_ */
