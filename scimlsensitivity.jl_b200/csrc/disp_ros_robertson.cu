// disp_ros_robertson.cu -- Rosenbrock23 kernels instantiated for the Robertson family
#include "disp_ros.inc"
namespace b200adj {
template int launch_ros_fwd<Robertson>(Handle*, const RosArgs&);
template int launch_ros_rev<Robertson>(Handle*, const RosArgs&);
}
