// micropp_geometry.cpp -- which material each hex8 element belongs to, for the 13 micro-structures.
//
// Host-only set-up code run once per micropp<3> object; the result (elem_type[nelem], values 0..2)
// is uploaded to the device and must be BIT-IDENTICAL to the reference's, so the
// floating-point geometry below evaluates the same expressions as src/micropp.cpp:339-544 on the
// same centroid coordinates (ex*dx + dx/2, ...), with the predicates of include/util.hpp.
// The numeric tables (sphere centres / radii, fibre directions) are data of the reference's
// micro-structure catalogue and have to match it digit for digit.
#include "micropp.hpp"

namespace {

// MIC3D_SPHERES: 40 inclusions, {cx, cy, cz, radius/0.1}
const double kSpheres[40][4] = {
    {.8663, .0689, .1568, .5741}, {.2305, .4008, .2093, .1735}, {.1987, .8423, .2126, .5065},
    {.6465, .7095, .4446, .8565}, {.6151, .9673, .2257, .9735}, {.1311, .9739, .4129, .7087},
    {.8433, .0738, .2233, .9585}, {.5124, .2111, .0369, .2843}, {.4094, .7030, .6241, .4029},
    {.1215, .8289, .3812, .2574}, {.1125, .9266, .6872, .1575}, {.7422, .4030, .6000, .3316},
    {.0791, .3819, .9297, .1874}, {.1201, .7712, .0069, .6300}, {.9339, .2177, .1976, .7049},
    {.3880, .1331, .9032, .7258}, {.7319, .7146, .8150, .8002}, {.8796, .2858, .6701, .8176},
    {.1427, .9309, .9830, .9970}, {.1388, .3126, .8054, .4736}, {.0729, .4754, .4950, .6351},
    {.5038, .7755, .7333, .3140}, {.0242, .2185, .9884, .3440}, {.5371, .5372, .8269, .2444},
    {.4564, .0039, .1715, .3894}, {.7410, .0799, .8775, .5234}, {.2627, .9047, .5681, .1658},
    {.7894, .1908, .2993, .4020}, {.1373, .5370, .9916, .6813}, {.0207, .8020, .9283, .2060},
    {.6540, .2359, .0286, .1427}, {.7801, .3568, .5086, .9332}, {.7322, .5706, .1114, .5406},
    {.5479, .1493, .1267, .9765}, {.6722, .1530, .1003, .0956}, {.7659, .3426, .9181, .6432},
    {.4582, .4636, .6310, .9998}, {.1811, .9665, .4713, .4166}, {.0834, .6066, .6936, .6907},
    {.4865, .7584, .3635, .0404}};

// MIC3D_FIBS_20_DISORDER: 20 fibres of radius 0.05, {dir_y, dir_z} (dir_x = 1); centres are the first
// 20 rows of kSpheres.
const double kFibreDir[20][2] = {{.5741, .8515}, {.1735, .1103}, {.5065, .4600}, {.8565, .9045}, {.9735, .6313},
                                 {.7087, .1547}, {.9585, .0220}, {.2843, .4062}, {.4029, .8095}, {.2574, .4742},
                                 {.1575, .0768}, {.3316, .0320}, {.1874, .6364}, {.6300, .2688}, {.7049, .5137},
                                 {.7258, .2799}, {.8002, .4794}, {.8176, .4142}, {.9970, .2189}, {.4736, .6202}};

const double kAxisX[3] = {1, 0, 0};
const double kAxisZ[3] = {0, 0, 1};

inline bool near2(double a, double ca, double b, double cb, double w) { return fabs(a - ca) < w && fabs(b - cb) < w; }

}  // namespace

// Free-function form (also used to classify the elements of one z-slab of a larger RVE): the unit cube
// lx = ly = lz = 1 of src/micropp.cpp:46-48.
int mpp_elem_type(int micro_type, const double *geo_params, double dx, double dy, double dz, int ex, int ey, int ez) {
  const double lx = 1.0, ly = 1.0, lz = 1.0;
  const double p[3] = {ex * dx + dx / 2., ey * dy + dy / 2., ez * dz + dz / 2.};
  const double mid[3] = {lx / 2, ly / 2, lz / 2};

  switch (micro_type) {
    case MIC_HOMOGENEOUS:
      return 0;

    case MIC_SPHERE:  // one centred inclusion of radius geo_params[0]
      return point_inside_sphere(mid, geo_params[0], p);

    case MIC_LAYER_Y:  // two flat layers, interface at y = geo_params[0]
      return (p[1] < geo_params[0]);

    case MIC_CILI_FIB_X:
      return point_inside_cilinder_inf(kAxisX, mid, geo_params[0], p);

    case MIC_CILI_FIB_Z:
      return point_inside_cilinder_inf(kAxisZ, mid, geo_params[0], p);

    case MIC_CILI_FIB_XZ: {  // a z fibre through y=0.75 and an x fibre through y=0.25
      const double c1[3] = {lx / 2., ly * .75, lz / 2.};
      const double c2[3] = {lx / 2., ly * .25, lz / 2.};
      return (point_inside_cilinder_inf(kAxisZ, c1, geo_params[0], p) ||
              point_inside_cilinder_inf(kAxisX, c2, geo_params[0], p))
                 ? 1
                 : 0;
    }

    case MIC_QUAD_FIB_XYZ: {  // three square fibres along z, y and x
      const double w = geo_params[0];
      if (near2(p[0], mid[0], p[1], mid[1], w)) return 1;
      if (near2(p[0], mid[0], p[2], mid[2], w)) return 1;
      if (near2(p[1], mid[1], p[2], mid[2], w)) return 1;
      return 0;
    }

    case MIC_QUAD_FIB_XZ: {
      const double w = geo_params[0];
      if (near2(p[0], mid[0], p[1], mid[1], w)) return 1;
      if (near2(p[1], mid[1], p[2], mid[2], w)) return 1;
      return 0;
    }

    case MIC_QUAD_FIB_XZ_BROKEN_X: {  // as above with a gap in the x fibre for 0.8 <= x <= 0.9
      const double w = geo_params[0];
      if (near2(p[0], mid[0], p[1], mid[1], w)) return 1;
      if (near2(p[1], mid[1], p[2], mid[2], w) && (p[0] < lx * .8 || p[0] > lx * .9)) return 1;
      return 0;
    }

    case MIC3D_SPHERES:
      for (int i = 0; i < 40; ++i) {
        const double r = 0.1 * kSpheres[i][3];
        if (point_inside_sphere(kSpheres[i], r, p)) return 1;
      }
      return 0;

    case MIC3D_8: {
      // 0 matrix, 1 the four cylinders, 2 their coating and the flat mid layer
      const double rad = 0.1, flat = 0.02, coat = 0.01;
      const double c1[3] = {lx * .25, ly * .75, 0.0};
      const double c2[3] = {lx * .75, ly * .75, 0.0};
      const double c3[3] = {0.0, ly * .25, lz * .25};
      const double c4[3] = {0.0, ly * .25, lz * .75};
      if (point_inside_cilinder_inf(kAxisZ, c1, rad, p) || point_inside_cilinder_inf(kAxisZ, c2, rad, p) ||
          point_inside_cilinder_inf(kAxisX, c3, rad, p) || point_inside_cilinder_inf(kAxisX, c4, rad, p))
        return 1;
      if (point_inside_cilinder_inf(kAxisZ, c1, rad + coat, p) ||
          point_inside_cilinder_inf(kAxisZ, c2, rad + coat, p) ||
          point_inside_cilinder_inf(kAxisX, c3, rad + coat, p) ||
          point_inside_cilinder_inf(kAxisX, c4, rad + coat, p) || fabs(p[1] - ly / 2) < flat)
        return 2;
      return 0;
    }

    case MIC3D_FIBS_20_ORDER: {  // 5 x 4 regular array of x fibres
      const double radius = 0.05;
      const int fibs_z = 5, fibs_y = 4;
      const double sz = 1.0 / (fibs_z + 1), sy = 1.0 / (fibs_y + 1);
      for (int i = 0; i < fibs_z; ++i)
        for (int j = 0; j < fibs_y; ++j) {
          const double c[3] = {0.0, (j + 1) * sy, (i + 1) * sz};
          if (point_inside_cilinder_inf(kAxisX, c, radius, p)) return 1;
        }
      return 0;
    }

    case MIC3D_FIBS_20_DISORDER: {
      const double radius = 0.05;
      for (int i = 0; i < 20; ++i) {
        const double dir[3] = {1, kFibreDir[i][0], kFibreDir[i][1]};
        if (point_inside_cilinder_inf(dir, kSpheres[i], radius, p)) return 1;
      }
      return 0;
    }
  }

  cerr << "Invalid micro_type = " << micro_type << endl;
  return -1;
}

template <>
int micropp<3>::get_elem_type(int ex, int ey, int ez) const {
  return mpp_elem_type(micro_type, geo_params, dx, dy, dz, ex, ey, ez);
}
