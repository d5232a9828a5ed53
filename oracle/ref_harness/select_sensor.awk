# Selects the lidar block of the reference's utility.h (UT:62-84) the way its README tells users to: the VLP-16 lines
# get commented out, the lines of the wanted sensor ("hdl32e" or "vls128") lose their "// ".  A block runs from its
# "// <name>" title to the next empty line.
BEGIN { title["hdl32e"] = "// HDL-32E"; title["vls128"] = "// VLS-128"; mode = 0 }
{
    if ($0 == "// VLP-16") { mode = 1; print; next }
    if ($0 == title[want]) { mode = 2; print; next }
    if ($0 ~ /^[ \t]*$/) { mode = 0; print; next }
    if (mode == 1) { print "// " $0; next }
    if (mode == 2) { sub(/^\/\/ /, ""); print; next }
    print
}
