/*
 * contour3d.h -- the drop-in boundary.
 *
 * Source-compatible restatement of the reference's only public header
 * (NaokiHori/Contour3D include/contour3d.h:1-88): the same four types with the
 * same member names, order and layout (sizeof contour3d_contour_obj_t = 88,
 * contour3d_line_obj_t = 80, contour3d_color_t = 3 on x86-64 SysV) and the
 * same single entry point, so a solver that links the reference can link
 * libcontour3d_b200.so instead without touching its code.
 *
 * Implemented in contour3d_b200/csrc/contour3d_execute.c (host C) on top of
 * the CUDA C-ABI of contour3d_b200.h.  There is no CPU fallback: without a
 * usable CUDA device the call logs "[CONTOUR3D ERROR] ..." and returns 1.
 *
 * The interface restated here is that of NaokiHori/Contour3D, distributed under
 * the MIT License, Copyright (c) 2023 NaokiHori (the permission notice is
 * reproduced in include/sdecomp.h, which restates the same author's SimpleDecomp
 * interface under the same licence).
 */
#if !defined(CONTOUR3D_H)
#define CONTOUR3D_H

#include <stddef.h>
#include <stdint.h>
#include "sdecomp.h"

#define CONTOUR3D_NDIMS 3

/* a point / direction in three dimensions */
typedef struct {
  double x;
  double y;
  double z;
} contour3d_vector_t;

/* 8-bit RGB */
typedef struct {
  uint8_t r;
  uint8_t g;
  uint8_t b;
} contour3d_color_t;

/* one iso-surface (reference: include/contour3d.h:22-40) */
typedef struct {
  sdecomp_pencil_t pencil;                 /* pencil the array lives on (X1 / X2: x fastest) */
  size_t glsizes[CONTOUR3D_NDIMS];         /* global number of grid points */
  double * grids[CONTOUR3D_NDIMS];         /* global rectilinear coordinates, glsizes[d] each */
  contour3d_vector_t (* converter) (       /* orthogonal -> Cartesian; called on the host, tabulated for the device */
      const contour3d_vector_t orthogonal
  );
  double threshold;                        /* iso value */
  contour3d_color_t color;                 /* surface colour */
  double * array;                          /* local pencil, x fastest; host OR device pointer */
} contour3d_contour_obj_t;

/* one poly-line (reference: include/contour3d.h:42-56) */
typedef struct {
  contour3d_vector_t edges[2];             /* first and last point, orthogonal coordinates */
  size_t nitems;                           /* number of points (nitems - 1 segments) */
  contour3d_vector_t (* converter) (
      const contour3d_vector_t orthogonal
  );
  contour3d_color_t color;
  double width;                            /* in pixels */
} contour3d_line_obj_t;

/* render one frame and write it as a binary PPM (reference: include/contour3d.h:58-86,
   src/contour3d/main.c:68-146).  Collective over the communicator held by
   sdecomp_info.  Returns 0 on success, 1 on failure. */
extern int contour3d_execute (
    const sdecomp_info_t * const sdecomp_info,
    const contour3d_vector_t * const camera_position,
    const contour3d_vector_t * const camera_look_at,
    const contour3d_vector_t * const light_direction,
    const size_t screen_sizes[2],
    const contour3d_vector_t * const screen_center,
    const contour3d_vector_t screen_local[2],
    const contour3d_color_t * const bg_color,
    const size_t num_contours,
    const contour3d_contour_obj_t * const contour3d_contour_objs,
    const size_t num_lines,
    const contour3d_line_obj_t * const contour3d_line_objs,
    const char fname[]
);

#endif /* CONTOUR3D_H */
