// ORACLE — TEST INFRASTRUCTURE ONLY. Stand-in for <ros/package.h> (included by the reference's map_manager/GridMap3D.h, nothing of it used).
#pragma once
