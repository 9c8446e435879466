#pragma once
#include <pluginlib/class_list_macros.h>
