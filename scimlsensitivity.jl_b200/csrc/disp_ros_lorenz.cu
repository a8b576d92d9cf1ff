// disp_ros_lorenz.cu -- Rosenbrock23 kernels instantiated for the Lorenz family
#include "disp_ros.inc"
namespace b200adj {
template int launch_ros_fwd<Lorenz>(Handle*, const RosArgs&);
template int launch_ros_rev<Lorenz>(Handle*, const RosArgs&);
}
