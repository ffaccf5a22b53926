/* stand-in for VTK's vtkUnsignedCharArray.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
