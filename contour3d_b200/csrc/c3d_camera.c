/*
 * c3d_camera.c -- camera / screen vectors from viewing angles (host only).
 *
 * The reference ships a browser tool for picking the arguments of
 * contour3d_execute (CameraAdjuster/src/config.ts:140-192): the camera sits
 * at distance d_camera from the origin on the z axis, tilted by `elevation`
 * about x and turned by `azimuth` about z; the screen's horizontal and vertical
 * vectors get the same two rotations plus a `roll` about the screen normal;
 * the screen centre lies on the line origin -> camera at distance d_screen.
 * This is the same construction as a C helper, so that benchmark and test
 * cameras (e.g. 16:9 screens) can be generated without the web page.
 */
#include <math.h>
#include <stddef.h>
#include "contour3d_b200.h"

/* rotation of b by angle t about the unit axis n (Rodrigues), CameraAdjuster/src/vector.ts:39-62 */
static void rotate(const double n[3], const double t, double b[3]) {
  const double c = cos(t), s = sin(t), k = 1. - c;
  const double r[3] = {
    (n[0] * n[0] * k + c) * b[0] + (n[0] * n[1] * k - n[2] * s) * b[1] + (n[0] * n[2] * k + n[1] * s) * b[2],
    (n[1] * n[0] * k + n[2] * s) * b[0] + (n[1] * n[1] * k + c) * b[1] + (n[1] * n[2] * k - n[0] * s) * b[2],
    (n[2] * n[0] * k - n[1] * s) * b[0] + (n[2] * n[1] * k + n[0] * s) * b[1] + (n[2] * n[2] * k + c) * b[2],
  };
  b[0] = r[0]; b[1] = r[1]; b[2] = r[2];
}

int c3d_camera_from_angles(double elevation_deg, double azimuth_deg, double roll_deg, double camera_distance, double screen_distance,
                           double screen_width, double screen_height, c3d_scene_t *scene) {
  if (NULL == scene || !(camera_distance > 0.) || !(screen_distance > 0.) || !(screen_width > 0.) || !(screen_height > 0.)) return 1;
  const double rad = acos(-1.) / 180.;
  const double ex[3] = {1., 0., 0.}, ez[3] = {0., 0., 1.};
  double cam[3] = {0., 0., camera_distance};
  double hor[3] = {screen_width, 0., 0.}, ver[3] = {0., screen_height, 0.};
  rotate(ex, elevation_deg * rad, cam);
  rotate(ez, azimuth_deg * rad, cam);
  rotate(ex, elevation_deg * rad, hor);
  rotate(ex, elevation_deg * rad, ver);
  rotate(ez, azimuth_deg * rad, hor);
  rotate(ez, azimuth_deg * rad, ver);
  double nrm[3] = {hor[1] * ver[2] - hor[2] * ver[1], hor[2] * ver[0] - hor[0] * ver[2], hor[0] * ver[1] - hor[1] * ver[0]};
  const double len = sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
  for (int d = 0; d < 3; d++) nrm[d] /= len;
  rotate(nrm, roll_deg * rad, hor);
  rotate(nrm, roll_deg * rad, ver);
  const double rate = screen_distance / camera_distance;
  for (int d = 0; d < 3; d++) {
    scene->camera[d] = cam[d];
    scene->center[d] = rate * cam[d];
    scene->local_x[d] = hor[d];
    scene->local_y[d] = ver[d];
  }
  return 0;
}
