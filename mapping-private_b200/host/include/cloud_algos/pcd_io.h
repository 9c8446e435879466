// pcd_io.h -- the little PCD I/O the GRSD tools need (color_chlac/include/color_chlac/grsd_colorCHLAC_tools.hpp:12-58):
//   readPCDXYZ    stands in for readPoints() -> pcl::io::loadPCDFile for the fields this path reads: x, y, z and, when
//                 present, normal_x, normal_y, normal_z and rgb (PCL's packed colour: the bits of a float / uint32 field,
//                 returned as 0x00RRGGBB).  Reads `DATA ascii` and `DATA binary` (v.7: the payload follows
//                 the DATA line; v.5/.6 files whose header is padded to a 4096-byte page, as the reference's
//                 color_chlac/demos/shape_data/*.pcd are: the payload then starts at byte 4096).
//   writeFeature  the reference's histogram writer: `FIELDS vfh`, COUNT = dimension, one ASCII line per histogram,
//                 all-zero histograms dropped when remove_0_flg is set (:32-58), values printed with "%f ".
// Unlike the reference's readPoints, which returns bool(-1) == true on failure (:16-19), a failed read returns false.
#ifndef CLOUD_ALGOS_PCD_IO_H
#define CLOUD_ALGOS_PCD_IO_H
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace cloud_algos
{

inline bool readPCDXYZ (const char* name, std::vector<float>& xyz, std::vector<float>* normals = 0, std::string* error = 0,
                        std::vector<uint32_t>* rgb = 0)
{
  xyz.clear ();
  if (normals) normals->clear ();
  if (rgb) rgb->clear ();
  std::ifstream fs (name, std::ios::binary);
  if (!fs.is_open ()) { if (error) *error = std::string ("Couldn't read file ") + name; return false; }
  std::vector<std::string> fields;
  std::vector<int> size, count;
  std::vector<char> type;
  long points = -1, width = -1, height = 1;
  std::string data_mode, line;
  std::streampos payload = 0;
  while (std::getline (fs, line))
  {
    if (!line.empty () && line[line.size () - 1] == '\r') line.erase (line.size () - 1);
    std::istringstream is (line);
    std::string key;
    if (!(is >> key) || key[0] == '#') continue;
    std::string v;
    if (key == "FIELDS" || key == "COLUMNS") while (is >> v) fields.push_back (v);
    else if (key == "SIZE") while (is >> v) size.push_back (std::atoi (v.c_str ()));
    else if (key == "TYPE") while (is >> v) type.push_back (v[0]);
    else if (key == "COUNT") while (is >> v) count.push_back (std::atoi (v.c_str ()));
    else if (key == "WIDTH") is >> width;
    else if (key == "HEIGHT") is >> height;
    else if (key == "POINTS") is >> points;
    else if (key == "DATA") { is >> data_mode; payload = fs.tellg (); break; }
  }
  if (fields.empty () || data_mode.empty ()) { if (error) *error = "not a PCD file"; return false; }
  // the header of an untrusted file decides every allocation below: check it first
  if (points < 0)
  {
    if (width < 0 || height < 0 || (height > 0 && width > (1L << 31) / height))
      { if (error) *error = "PCD header without a usable POINTS / WIDTH x HEIGHT"; return false; }
    points = width * height;
  }
  if (points > (1L << 31) - 16) { if (error) *error = "PCD header declares too many points"; return false; }
  if (size.empty ()) size.assign (fields.size (), 4);   // old headers without SIZE / TYPE: float32 columns
  if (type.empty ()) type.assign (fields.size (), 'F');
  if (count.empty ()) count.assign (fields.size (), 1);
  if (size.size () != fields.size () || type.size () != fields.size () || count.size () != fields.size ())
    { if (error) *error = "inconsistent PCD header"; return false; }
  for (size_t f = 0; f < fields.size (); ++f)
    if ((size[f] != 1 && size[f] != 2 && size[f] != 4 && size[f] != 8) || count[f] < 1 || count[f] > 65536)
      { if (error) *error = "PCD header with an invalid SIZE or COUNT"; return false; }
  int col[6] = {-1, -1, -1, -1, -1, -1};   // column (in scalar elements) of x y z normal_x normal_y normal_z
  int off[6] = {-1, -1, -1, -1, -1, -1};   // byte offset inside a binary record
  int ncols = 0, rec = 0;
  int rgb_col = -1, rgb_off = -1;
  char rgb_type = 'F';
  static const char* want[6] = {"x", "y", "z", "normal_x", "normal_y", "normal_z"};
  for (size_t f = 0; f < fields.size (); ++f)
  {
    if ((fields[f] == "rgb" || fields[f] == "rgba") && size[f] == 4) { rgb_col = ncols; rgb_off = rec; rgb_type = type[f]; }
    for (int w = 0; w < 6; ++w)
      if (fields[f] == want[w] && size[f] == 4 && type[f] == 'F') { col[w] = ncols; off[w] = rec; }
    ncols += count[f];
    rec += size[f] * count[f];
  }
  if (col[0] < 0 || col[1] < 0 || col[2] < 0) { if (error) *error = "PCD file without float x, y, z fields"; return false; }
  const bool have_n = normals && col[3] >= 0 && col[4] >= 0 && col[5] >= 0;
  const bool have_rgb = rgb && rgb_col >= 0;
  {  // the payload must be there before anything of its declared size is allocated (an ASCII row has >= 2 bytes)
    const std::streampos here = fs.tellg ();
    fs.seekg (0, std::ios::end);
    const std::streamoff have = fs.tellg ();
    fs.seekg (here);
    const std::streamoff per_point = data_mode == "binary" ? (std::streamoff) rec : 2;
    if (per_point <= 0 || (std::streamoff) points > have / per_point) { if (error) *error = "truncated PCD file"; return false; }
  }
  xyz.resize ((size_t) points * 3);
  if (have_n) normals->resize ((size_t) points * 3);
  if (have_rgb) rgb->resize ((size_t) points);
  if (data_mode == "ascii")
  {
    std::vector<double> row (ncols);
    for (long p = 0; p < points; ++p)
    {
      if (!std::getline (fs, line)) { if (error) *error = "truncated PCD file"; return false; }
      std::istringstream is (line);
      for (int c = 0; c < ncols; ++c)
      {
        std::string tok;
        if (!(is >> tok)) { if (error) *error = "short PCD row"; return false; }
        row[c] = std::strtod (tok.c_str (), 0);   // "nan" parses as NaN
      }
      for (int a = 0; a < 3; ++a) xyz[3 * (size_t) p + a] = (float) row[col[a]];
      if (have_n) for (int a = 0; a < 3; ++a) (*normals)[3 * (size_t) p + a] = (float) row[col[3 + a]];
      if (have_rgb)
      {
        uint32_t bits;
        if (rgb_type == 'F') { const float f = (float) row[rgb_col]; std::memcpy (&bits, &f, 4); }   // PCL prints the float whose bits are the colour
        else bits = (uint32_t) row[rgb_col];
        (*rgb)[(size_t) p] = bits & 0x00ffffffu;
      }
    }
    return true;
  }
  if (data_mode != "binary") { if (error) *error = "unsupported PCD DATA mode " + data_mode; return false; }
  fs.seekg (0, std::ios::end);
  const std::streamoff file_size = fs.tellg ();
  if (rec <= 0) { if (error) *error = "PCD header with empty records"; return false; }
  const std::streamoff need = (std::streamoff) points * rec;
  // page-padded header of the older writers: exactly one 4096-byte page before the records
  if (file_size - 4096 == need && (std::streamoff) payload <= 4096) payload = 4096;
  if (file_size - (std::streamoff) payload < need) { if (error) *error = "truncated PCD file"; return false; }
  fs.seekg (payload);
  std::vector<char> buf ((size_t) need);
  fs.read (buf.data (), need);
  for (long p = 0; p < points; ++p)
  {
    const char* r = buf.data () + (size_t) p * rec;
    for (int a = 0; a < 3; ++a) std::memcpy (&xyz[3 * (size_t) p + a], r + off[a], 4);
    if (have_n) for (int a = 0; a < 3; ++a) std::memcpy (&(*normals)[3 * (size_t) p + a], r + off[3 + a], 4);
    if (have_rgb) { uint32_t bits; std::memcpy (&bits, r + rgb_off, 4); (*rgb)[(size_t) p] = bits & 0x00ffffffu; }
  }
  return true;
}

inline bool if_zero_vec (const std::vector<float>& vec)
{
  for (size_t i = 0; i < vec.size (); i++)
    if (vec[i] != 0) return false;
  return true;
}

inline bool writeFeature (const char* name, const std::vector< std::vector<float> >& feature, bool remove_0_flg = true)
{
  if (feature.empty ()) return false;
  const int feature_size = (int) feature[0].size ();
  const int hist_num_init = (int) feature.size ();
  int hist_num = hist_num_init;
  if (remove_0_flg)
    for (int h = 0; h < hist_num_init; h++)
      if (if_zero_vec (feature[h])) hist_num--;
  FILE* fp = std::fopen (name, "w");
  if (!fp) return false;
  std::fprintf (fp, "# .PCD v.7 - Point Cloud Data file format\n");
  std::fprintf (fp, "FIELDS vfh\n");
  std::fprintf (fp, "SIZE 4\n");
  std::fprintf (fp, "TYPE F\n");
  std::fprintf (fp, "COUNT %d\n", feature_size);
  std::fprintf (fp, "WIDTH %d\n", hist_num);
  std::fprintf (fp, "HEIGHT 1\n");
  std::fprintf (fp, "POINTS %d\n", hist_num);
  std::fprintf (fp, "DATA ascii\n");
  for (int h = 0; h < hist_num_init; h++)
  {
    if (remove_0_flg && if_zero_vec (feature[h])) continue;
    for (int t = 0; t < feature_size; t++) std::fprintf (fp, "%f ", feature[h][t]);
    std::fprintf (fp, "\n");
  }
  std::fclose (fp);
  return true;
}

}
#endif
