// disp_ros_lv.cu -- Rosenbrock23 kernels instantiated for the LotkaVolterra family
#include "disp_ros.inc"
namespace b200adj {
template int launch_ros_fwd<LotkaVolterra>(Handle*, const RosArgs&);
template int launch_ros_rev<LotkaVolterra>(Handle*, const RosArgs&);
}
