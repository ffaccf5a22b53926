/* stand-in for VTK's vtkMatrix3x3.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
