/* stand-in for VTK's vtkIntArray.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
