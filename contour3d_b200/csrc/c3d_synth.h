/* constants shared by the device generator (k_synth, c3d_kernels.cu) and its host twin (c3d_synth_host.c) */
#if !defined(C3D_SYNTH_H)
#define C3D_SYNTH_H
/* kind 2: value = C3D_CURL_SCALE * |curl F| - 1 (the median of |curl F| of a 5-octave field is about 1 / C3D_CURL_SCALE) */
#define C3D_CURL_SCALE 0.18f
#endif
