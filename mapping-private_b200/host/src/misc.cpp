// misc.cpp -- helpers shared by the plugins (reference: cloud_algos/src/misc.cpp:4-25).
#include <cloud_algos/cloud_algos.h>

#include "cloud_algos_b200.h"

namespace cloud_algos
{

int getChannelIndex (const sensor_msgs::PointCloud &points, std::string value)
{
  for (size_t d = 0; d < points.channels.size (); ++d)
    if (points.channels[d].name == value) return (int) d;
  return -1;
}

int getChannelIndex (const boost::shared_ptr<const sensor_msgs::PointCloud> points, std::string value)
{
  return getChannelIndex (*points, value);
}

GpuContext::~GpuContext ()
{
  if (ctx_) cab_destroy (ctx_);
}

cab_ctx* GpuContext::get (std::string &error, bool exact)
{
  if (ctx_ && exact_ == exact) return ctx_;
  if (ctx_) { cab_destroy (ctx_); ctx_ = 0; }
  cab_config cfg = {};
  cfg.device = 0;
  cfg.exact = exact ? 1 : 0;
  if (cab_create (&cfg, &ctx_) != CAB_OK)
  {
    error = std::string ("GPU unavailable: ") + cab_last_error (0);
    ctx_ = 0;
    return 0;
  }
  exact_ = exact;
  return ctx_;
}

}
