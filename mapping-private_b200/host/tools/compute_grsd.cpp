// compute_grsd -- the B200 counterpart of the reference's feature extraction tools
// (color_feature_classification/test/computeGRSD.cpp:22-47,181-230 and color_chlac/test/exampleRSD.cpp:50-93):
// read a PCD file, estimate normals with a 2 cm radius unless the file carries them, voxelise, compute the GRSD
// signature(s) -- one per cluster, or one per sliding box when -subdiv is given, repeated over the voxel offsets
// 0, step, 2*step, ... < subdiv on every axis -- and write them with the reference's writeFeature format.
//
//   compute_grsd input.pcd voxel_size output.pcd [-subdiv N] [-offset n] [-kind 21|325|110|vosch|c3hlac|cchlac] [-normalize]
//
// -kind vosch writes what color_chlac/test/exampleVOSCH.cpp and color_feature_classification/test/computeVOSCH.cpp write:
// the 20 GRSD-21 values followed by the 117 rotation-invariant C3-HLAC values of the voxel colours (thresholds 127);
// c3hlac / cchlac write the colour part alone (C3-HLAC, or the Color-CHLAC coding of example_GRSD_CCHLAC.cpp's files).
// Differences to computeGRSD.cpp: the voxel size is an argument instead of {config_txt_path}/voxel_size.txt, the
// rotation augmentation (-rotate) is not offered, and -kind selects the other signatures as well.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cloud_algos/pcd_io.h>

#include "cloud_algos_b200.h"

static int parse_int (int argc, char** argv, const char* flag, int& value)
{
  for (int i = 1; i + 1 < argc; ++i)
    if (std::strcmp (argv[i], flag) == 0) { value = std::atoi (argv[i + 1]); return i; }
  return -1;
}

int main (int argc, char** argv)
{
  if (argc < 4)
  {
    std::fprintf (stderr, "Need at least three parameters! Syntax is: %s {input_pointcloud_filename.pcd} {voxel_size} "
                          "{output_histogram_filename.pcd} [options]\n"
                          "                          -subdiv N = subdivision size (e.g. 10 voxels)\n"
                          "                          -offset n = offset step for subdivisions (e.g. 5 voxels)\n"
                          "                          -kind 21|325|110 = GRSD-21 (default), GRSD-325, PlusGRSD-110\n"
                          "                          -kind vosch|c3hlac|cchlac = VOSCH (GRSD-21 + C3-HLAC), C3-HLAC, Color-CHLAC (input needs rgb)\n"
                          "                          -normalize = multiply by NORMALIZE_GRSD = 20/26\n", argv[0]);
    return -1;
  }
  const float voxel_size = (float) std::atof (argv[2]);
  int subdivision_size = 0, offset_step = 1, kind_arg = 21;
  if (parse_int (argc, argv, "-subdiv", subdivision_size) > 0 && subdivision_size < 0)
  { std::fprintf (stderr, "Invalid subdivision size (%d)! \n", subdivision_size); return -1; }
  if (parse_int (argc, argv, "-offset", offset_step) > 0 && (offset_step < 1 || offset_step >= subdivision_size))
  { std::fprintf (stderr, "Invalid offset step (%d)! (while subdivision size is %d.)\n", offset_step, subdivision_size); return -1; }
  parse_int (argc, argv, "-kind", kind_arg);
  // colour modes: 0 none, 1 VOSCH (GRSD-21 + C3-HLAC), 2 C3-HLAC alone, 3 Color-CHLAC alone
  int color_mode = 0;
  for (int i = 4; i + 1 < argc; ++i)
    if (std::strcmp (argv[i], "-kind") == 0)
      color_mode = std::strcmp (argv[i + 1], "vosch") == 0 ? 1 : std::strcmp (argv[i + 1], "c3hlac") == 0 ? 2 : std::strcmp (argv[i + 1], "cchlac") == 0 ? 3 : 0;
  if (color_mode) kind_arg = 21;
  bool normalize = false;
  for (int i = 4; i < argc; ++i) if (std::strcmp (argv[i], "-normalize") == 0) normalize = true;
  const int kind = kind_arg == 325 ? CAB_SIG_GRSD325 : (kind_arg == 110 ? CAB_SIG_PLUSGRSD110 : CAB_SIG_GRSD21);
  const int dim = kind == CAB_SIG_GRSD325 ? 325 : (kind == CAB_SIG_PLUSGRSD110 ? 110 : 21);
  const int emit = kind == CAB_SIG_GRSD21 ? 20 : dim;   // GRSD-21 drops its (EMPTY, EMPTY) bin (grsd_colorCHLAC_tools.hpp:278-292)
  if (!(voxel_size > 0)) { std::fprintf (stderr, "Invalid voxel size %s\n", argv[2]); return -1; }

  std::vector<float> xyz, nrm;
  std::vector<uint32_t> rgb;
  std::string err;
  if (!cloud_algos::readPCDXYZ (argv[1], xyz, &nrm, &err, &rgb)) { std::fprintf (stderr, "%s\n", err.c_str ()); return -1; }
  const int n = (int) (xyz.size () / 3);
  if (n == 0) { std::fprintf (stderr, "%s holds no points\n", argv[1]); return -1; }
  if (color_mode && (int) rgb.size () != n) { std::fprintf (stderr, "%s has no rgb field\n", argv[1]); return -1; }

  cab_config cfg = {};
  cfg.exact = 1;
  cab_ctx* ctx = 0;
  if (cab_create (&cfg, &ctx) != CAB_OK) { std::fprintf (stderr, "GPU unavailable: %s\n", cab_last_error (0)); return -2; }
  const int32_t offsets[2] = {0, n};
  const float vp[3] = {0, 0, 0};
  int32_t hist21[21];
  std::vector<float> nx, ny, nz;
  if (!nrm.empty ())
  {
    nx.resize (n); ny.resize (n); nz.resize (n);
    for (int i = 0; i < n; ++i) { nx[i] = nrm[3 * i]; ny[i] = nrm[3 * i + 1]; nz[i] = nrm[3 * i + 2]; }
  }
  int rc = cab_grsd_batch (ctx, &xyz[0], 3, offsets, 1, voxel_size, /*normals_radius_search*/ 0.02f, /*rsd_radius_search*/ 0.01, 0, vp,
                           nrm.empty () ? 0 : &nx[0], nrm.empty () ? 0 : &ny[0], nrm.empty () ? 0 : &nz[0], hist21);
  if (rc != CAB_OK) { std::fprintf (stderr, "GRSD failed: %s\n", cab_last_error (ctx)); cab_destroy (ctx); return -2; }

  // repeat with changing offset values for subdivisions (computeGRSD.cpp:24-45)
  int repeat_num_offset = (int) std::ceil (subdivision_size / offset_step);   // integer division, as in the reference
  if (subdivision_size == 0) repeat_num_offset = 1;
  std::vector< std::vector<float> > feature;
  for (int ox = 0; ox < repeat_num_offset; ox++)
    for (int oy = 0; oy < repeat_num_offset; oy++)
      for (int oz = 0; oz < repeat_num_offset; oz++)
      {
        int64_t hoff[2] = {0, 0};
        const int64_t total = cab_grsd_signatures (ctx, kind, subdivision_size, ox * offset_step, oy * offset_step, oz * offset_step,
                                                   hoff, 0, 0, 0);
        if (total < 0) { std::fprintf (stderr, "GRSD failed: %s\n", cab_last_error (ctx)); cab_destroy (ctx); return -2; }
        if (total == 0) continue;
        std::vector<int32_t> h ((size_t) total * dim);
        if (cab_grsd_signatures (ctx, kind, subdivision_size, ox * offset_step, oy * offset_step, oz * offset_step, hoff, 0, &h[0], total) < 0)
        { std::fprintf (stderr, "GRSD failed: %s\n", cab_last_error (ctx)); cab_destroy (ctx); return -2; }
        std::vector<float> colour;
        if (color_mode)
        {
          colour.resize ((size_t) total * 117);
          const int64_t tc = cab_color_chlac (ctx, &rgb[0], color_mode == 3 ? 0 : 1, 127, 127, 127, subdivision_size, ox * offset_step,
                                              oy * offset_step, oz * offset_step, 0, 0, &colour[0], total);
          if (tc != total) { std::fprintf (stderr, "colour features failed: %s\n", cab_last_error (ctx)); cab_destroy (ctx); return -2; }
        }
        for (int64_t s = 0; s < total; ++s)
        {
          std::vector<float> f;
          if (color_mode <= 1)
            for (int i = 0; i < emit; ++i) f.push_back (h[(size_t) s * dim + i] * (normalize ? 20.0f / 26 : 1.0f));
          if (color_mode)   // conc_vector (grsd, c3_hlac), grsd_colorCHLAC_tools.hpp:823-843
            f.insert (f.end (), colour.begin () + s * 117, colour.begin () + (s + 1) * 117);
          feature.push_back (f);
        }
      }
  cab_destroy (ctx);
  if (feature.empty ()) { std::fprintf (stderr, "no histogram produced\n"); return -3; }
  if (!cloud_algos::writeFeature (argv[3], feature, subdivision_size > 0)) { std::fprintf (stderr, "Couldn't write %s\n", argv[3]); return -1; }
  std::printf ("%d points -> %d histogram(s) of %d values written to %s\n", n, (int) feature.size (), (int) feature[0].size (), argv[3]);
  return 0;
}
