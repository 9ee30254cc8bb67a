"""ros1_shim/ is a catkin package (package.xml + CMakeLists.txt) whose two sources compile and link: there is no
ROS in this image, so they are built against tests/ros_stubs/ -- the slice of the roscpp / sensor_msgs API they
use, with roscpp's signatures -- and against the real C ABI header and library.  CPU only."""
import os
import subprocess
import xml.etree.ElementTree as ET

import pytest

from conftest import ROOT

SHIM = os.path.join(ROOT, "ros1_shim")
NODES = ["disparity_to_point_cloud_node", "depth_map_fusion_node"]


def test_package_manifest_and_cmake_targets():
    pkg = ET.parse(os.path.join(SHIM, "package.xml")).getroot()
    assert pkg.findtext("name") == "disparity_to_point_cloud"       # the reference's package name (package.xml:3)
    assert pkg.findtext("buildtool_depend") == "catkin"
    build = {e.text for e in pkg.findall("build_depend")}
    run = {e.text for e in pkg.findall("run_depend")}
    assert {"roscpp", "sensor_msgs"} <= build and {"roscpp", "sensor_msgs"} <= run
    assert not ({"pcl_ros", "cv_bridge", "libpcl-all-dev"} & build)  # what the GPU library replaces
    cm = open(os.path.join(SHIM, "CMakeLists.txt")).read()
    assert "project(disparity_to_point_cloud)" in cm and "find_package(catkin REQUIRED COMPONENTS" in cm
    for node in NODES:                                                # reference CMakeLists.txt:149-150 executable names
        assert node in cm and os.path.exists(os.path.join(SHIM, node + ".cpp"))
    for launch, node in (("d2pcloud.launch", NODES[0]), ("depth_map_fusion.launch", NODES[1])):
        text = open(os.path.join(ROOT, "launch", launch)).read()
        assert 'pkg="disparity_to_point_cloud"' in text and f'type="{node}"' in text


@pytest.mark.parametrize("node", NODES)
def test_shim_compiles_and_links_against_the_c_abi(node, tmp_path):
    from disparity_to_point_cloud_b200 import build
    lib = build.build()
    exe = tmp_path / node
    cmd = ["g++", "-std=c++14", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "ros_stubs"), "-I",
           os.path.join(ROOT, "include"), os.path.join(SHIM, node + ".cpp"), "-o", str(exe), "-L", os.path.dirname(lib),
           "-ld2pc_b200", "-Wl,-rpath," + os.path.dirname(lib)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert exe.exists()
