#pragma once
#include "/root/reference/include/colour/aces.h"
#include "/root/reference/include/colour/adobergb.h"
#include "/root/reference/include/colour/ergb.h"
#include "/root/reference/include/colour/rec709.h"
#include "/root/reference/include/colour/srgb.h"
#include "/root/reference/include/colour/xyz.h"
#define colour_input_to_xyz colour_ergb_to_xyz
#define colour_xyz_to_input colour_xyz_to_ergb
#define colour_input_print_info colour_ergb_print_info
#define colour_output_to_xyz colour_srgb_to_xyz
#define colour_xyz_to_output colour_xyz_to_srgb
#define colour_output_print_info colour_srgb_print_info
#define colour_camera_to_xyz colour_xyz_to_xyz
#define colour_xyz_to_camera colour_xyz_to_xyz
#define colour_camera_print_info colour_xyz_print_info
