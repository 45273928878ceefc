// Inter-process hand-off: packing one message -- the tensors of a keyframe / frame tuple -- into a
// persistent device slot in ONE launch, with the dtype conversion of the reference's transfer_data
// (como/utils/multiprocessing.py:16-21: every tensor .to(device, dtype)) fused into the copy.  The slot ring, the
// CUDA IPC events and the control queue live in como_b200/utils/multiprocessing.py.
#include "common.cuh"

namespace como {

struct PackArgs {
  como_b200_pack_item_t item[COMO_B200_PACK_MAX_ITEMS];
  int n;
};

template <typename S, typename D>
__device__ __forceinline__ void pack_copy(const void* src, void* dst, long long count, long long start, long long stride) {
  const S* s = reinterpret_cast<const S*>(src);
  D* d = reinterpret_cast<D*>(dst);
  for (long long i = start; i < count; i += stride) d[i] = (D)s[i];
}

template <typename D>
__device__ __forceinline__ void pack_from(int src_dtype, const void* src, void* dst, long long count, long long start,
                                          long long stride) {
  switch (src_dtype) {
    case COMO_B200_DT_F32: pack_copy<float, D>(src, dst, count, start, stride); break;
    case COMO_B200_DT_F64: pack_copy<double, D>(src, dst, count, start, stride); break;
    case COMO_B200_DT_U8: pack_copy<uint8_t, D>(src, dst, count, start, stride); break;
    case COMO_B200_DT_I32: pack_copy<int32_t, D>(src, dst, count, start, stride); break;
    case COMO_B200_DT_I64: pack_copy<long long, D>(src, dst, count, start, stride); break;
    default: break;
  }
}

__global__ void __launch_bounds__(256) handoff_pack_kernel(PackArgs a, uint8_t* __restrict__ slot) {
  const int it = blockIdx.y;
  if (it >= a.n) return;
  const como_b200_pack_item_t& m = a.item[it];
  const long long start = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  void* dst = slot + m.dst_offset_bytes;
  if (m.dst_dtype == COMO_B200_DT_F32) pack_from<float>(m.src_dtype, m.src, dst, m.count, start, stride);
  else if (m.dst_dtype == COMO_B200_DT_F64) pack_from<double>(m.src_dtype, m.src, dst, m.count, start, stride);
}

}  // namespace como

using namespace como;

extern "C" int como_b200_handoff_pack(const como_b200_pack_item_t* items, int32_t num_items, void* slot, void* stream) {
  COMO_REQUIRE(items && slot, "handoff_pack: null pointer argument");
  COMO_REQUIRE(num_items >= 1 && num_items <= COMO_B200_PACK_MAX_ITEMS, "handoff_pack: 1..%d items per launch",
               COMO_B200_PACK_MAX_ITEMS);
  PackArgs a;
  a.n = num_items;
  long long most = 1;
  for (int i = 0; i < num_items; ++i) {
    const como_b200_pack_item_t& m = items[i];
    COMO_REQUIRE(m.src != nullptr && m.count >= 0 && m.dst_offset_bytes >= 0, "handoff_pack: bad item %d", i);
    COMO_REQUIRE(m.src_dtype >= COMO_B200_DT_F32 && m.src_dtype <= COMO_B200_DT_I64, "handoff_pack: item %d: source dtype", i);
    COMO_REQUIRE(m.dst_dtype == COMO_B200_DT_F32 || m.dst_dtype == COMO_B200_DT_F64, "handoff_pack: item %d: target dtype must be f32/f64", i);
    COMO_REQUIRE((m.dst_offset_bytes & 15) == 0, "handoff_pack: item %d: slot offsets are 16-byte aligned", i);
    a.item[i] = m;
    if (m.count > most) most = m.count;
  }
  long long bx = (most + 256 * 8 - 1) / (256 * 8);
  const long long cap = 4LL * sm_count();
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  handoff_pack_kernel<<<dim3((unsigned)bx, (unsigned)num_items), 256, 0, (cudaStream_t)stream>>>(a, (uint8_t*)slot);
  return check_launch("handoff_pack");
}
