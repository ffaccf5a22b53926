/* stand-in for VTK's vtkInformation.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
