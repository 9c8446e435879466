// cloud_algos.h -- the CloudAlgo plugin surface, kept name-for-name so that callers of the
// reference plugins compile and run against the B200 implementation unchanged.
// Mirrors cloud_algos/include/cloud_algos/cloud_algos.h:21-117 of the reference: the abstract base
// (virtual init/pre/post/requires/provides/createPublisher, public verbosity_level_ and
// output_valid_), the CloudAlgoNode<algo> subscriber wrapper and standalone_node<algo>.
// process() and output() are deliberately NOT virtual (cloud_algos.h:41-42): callers down-cast
// to the concrete class (dyn_obj_store/src/table_memory_grsd.cpp:975-987).
#ifndef CLOUD_ALGOS_H
#define CLOUD_ALGOS_H

#include <climits>
#include <string>
#include <vector>

#include <ros/ros.h>
#include <ros/node_handle.h>
#include <sensor_msgs/PointCloud.h>

struct cab_ctx;  // C ABI handle of libcloudalgos_b200 (include/cloud_algos_b200.h)

namespace cloud_algos
{

// index of the channel called `value`, -1 if absent (cloud_algos/src/misc.cpp:4-25)
int getChannelIndex (const sensor_msgs::PointCloud &points, std::string value);
int getChannelIndex (const boost::shared_ptr<const sensor_msgs::PointCloud> points, std::string value);

class CloudAlgo
{
 public:
  int verbosity_level_;
  bool output_valid_;

  CloudAlgo () : verbosity_level_ (INT_MAX), output_valid_ (true) {}
  virtual ~CloudAlgo () {}

  typedef void OutputType;
  typedef sensor_msgs::PointCloud InputType;
  static std::string default_output_topic () {return std::string ("");}
  static std::string default_input_topic () {return std::string ("");}
  static std::string default_node_name () {return std::string ("");}

  virtual void init (ros::NodeHandle&) = 0;
  virtual void pre () = 0;
  virtual void post () = 0;
  virtual std::vector<std::string> requires () = 0;
  virtual std::vector<std::string> provides () = 0;
  virtual ros::Publisher createPublisher (ros::NodeHandle& nh) = 0;
};

// Owns the GPU context of one plugin instance (one CUDA stream + grow-only device arena, reused
// across process() calls; the reference rebuilds its kd-tree every call, radius_estimation.cpp:55).
// Creation is lazy; failure (no B200, no library) makes process() report an error and clear
// output_valid_ -- there is no CPU fallback.
class GpuContext
{
 public:
  GpuContext () : ctx_ (0), exact_ (false) {}
  ~GpuContext ();
  cab_ctx* get (std::string &error, bool exact = false);
 private:
  GpuContext (const GpuContext&);
  GpuContext& operator= (const GpuContext&);
  cab_ctx* ctx_;
  bool exact_;
};

// Subscriber -> pre, process, publish(output), post   (cloud_algos.h:46-104)
template <class algo>
  class CloudAlgoNode
{
 public:
  CloudAlgoNode (ros::NodeHandle& nh, algo &alg) : nh_ (nh), a (alg)
  {
    // check if the input topic is advertised (cloud_algos.h:53-66)
    std::vector<ros::master::TopicInfo> t_list;
    bool topic_found = false;
    ros::master::getTopics (t_list);
    for (std::vector<ros::master::TopicInfo>::iterator it = t_list.begin (); it != t_list.end (); it++)
      if (it->name == a.default_input_topic ()) { topic_found = true; break; }
    if (!topic_found)
      ROS_WARN ("Trying to subscribe to %s, but the topic doesn't exist!", a.default_input_topic ().c_str ());

    pub_ = nh_.advertise <typename algo::OutputType> (a.default_output_topic (), 5);
    sub_ = nh_.subscribe (a.default_input_topic (), 1, &CloudAlgoNode<algo>::input_cb, this);
    ROS_INFO ("CloudAlgoNode (%s) created. SUB [%s], PUB[%s]", ros::this_node::getName ().c_str (),
              a.default_input_topic ().c_str (), a.default_output_topic ().c_str ());
    a.init (nh_);
  }

  void input_cb (const boost::shared_ptr<const typename algo::InputType> &input)
  {
    a.pre ();
    std::string result = a.process (input);
    ROS_INFO ("Result got after processed message: %s", result.c_str ());
    if (a.output_valid_)
      pub_.publish (a.output ());
    else
      ROS_ERROR ("Not publishing result as it is invalid!");
    a.post ();
  }

  ros::NodeHandle& nh_;
  ros::Publisher pub_;
  ros::Subscriber sub_;
  algo& a;
};

// main() of the <algo>_node executables (cloud_algos.h:106-117; built with -DCREATE_NODE, CMakeLists.txt:42-57)
template <class algo>
  int standalone_node (int argc, char* argv[])
{
  ros::init (argc, argv, algo::default_node_name ());
  algo a;
  ros::NodeHandle nh ("~");
  CloudAlgoNode<algo> c (nh, a);
  ros::spin ();
  return (0);
}

}
#endif
