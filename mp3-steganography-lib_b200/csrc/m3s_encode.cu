// m3s_encode.cu -- encode half (placeholder while the decode half is being validated on hardware)
#include "m3s_common.cuh"

extern "C" int64_t m3s_encode_bound(int64_t n_samples, int32_t sample_rate, int32_t bitrate_kbps)
{
    if (n_samples < 0 || sample_rate <= 0) return -1;
    int64_t frames = (n_samples + 1151) / 1152;
    int64_t fs = (144000LL * bitrate_kbps) / sample_rate + 1;
    return frames * fs + 8;
}

extern "C" int m3s_encode(m3s_handle_t h, const int16_t *, int, const int64_t *, const int64_t *, int32_t, int32_t, int32_t,
                          const uint8_t *, const int64_t *, uint8_t *, const int64_t *, const int64_t *, int64_t *, int64_t *)
{
    return m3s_fail(h, M3S_ERR_STATE, "m3s_encode: not built yet");
}

extern "C" int m3s_encode_taps(m3s_handle_t h, int32_t *, int32_t *, int32_t *, int32_t *)
{
    return m3s_fail(h, M3S_ERR_STATE, "m3s_encode_taps: not built yet");
}
